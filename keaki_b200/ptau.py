"""snarkjs `.ptau` container reader and writer — host-side setup I/O (src/kzg/ptau.rs:12-376).

Same container rules as the reference parser (magic "ptau", 12-byte metadata, 11 sections with
12-byte headers, header section = n8, modulus, power, ceremony power; TauG1 has 2*2^power - 1 points,
TauG2 has 2^power).  DELIBERATE DEVIATION: snarkjs stores coordinates as
Montgomery limbs; the reference reads them as canonical integers (`deserialize_uncompressed_unchecked`,
src/kzg/ptau.rs:266,314) and silently gets off-curve points.  Since the C ABI *wants* Montgomery limbs,
this reader passes the file's limbs through unchanged — which is the correct decoding."""
from __future__ import annotations

import struct

import numpy as np

from .types import FQ_MODULUS

FILE_TYPE = b"ptau"
N_SECTIONS = 11
METADATA_LEN = 12
SECTION_HEADER_LEN = 12
SECTION_IDS = (1, 2, 3, 4, 5, 6, 7, 12, 13, 14, 15)  # src/kzg/ptau.rs:24-51


class SetupFileError(Exception):
    """src/kzg/ptau.rs:360-376"""


def parse_sections(data: bytes):
    if len(data) < METADATA_LEN or data[:4] != FILE_TYPE:
        raise SetupFileError("InvalidFileType")
    _version, n_sections = struct.unpack_from("<II", data, 4)
    if n_sections != N_SECTIONS:
        raise SetupFileError(f"InvalidSectionCount({n_sections})")
    off, sections = METADATA_LEN, {}
    for _ in range(n_sections):
        if off + SECTION_HEADER_LEN > len(data):
            raise SetupFileError("UnexpectedEof")
        sid, slen = struct.unpack_from("<IQ", data, off)
        if sid not in SECTION_IDS:
            raise SetupFileError(f"UnknownSection({sid})")
        off += SECTION_HEADER_LEN
        if off + slen > len(data):
            raise SetupFileError("UnexpectedEof")
        sections.setdefault(sid, (off, slen))   # a repeated id keeps its first occurrence
        off += slen
    if off != len(data):
        raise SetupFileError("SectionsNotContiguous")
    return sections


MAX_POWER = 28   # 2-adicity of the BN254 scalar field: no larger ceremony exists (and 2^power * 128 B must stay addressable)


def _section(sections, sid):
    """`FileSections::get` (src/kzg/ptau.rs:146-150): a section that is not in the file is EmptySection(id)."""
    if sid not in sections:
        raise SetupFileError(f"EmptySection({sid})")
    return sections[sid]


def read_header(data: bytes, sections):
    """`HeaderSection::parse` (src/kzg/ptau.rs:190-222); short headers are a ParseError instead of the reference's slice panic."""
    off, slen = _section(sections, 1)
    if slen < 4:
        raise SetupFileError("ParseError(header section too short)")
    n8 = struct.unpack_from("<I", data, off)[0]
    if 4 + n8 + 8 > slen:
        raise SetupFileError("ParseError(header section too short)")
    modulus = int.from_bytes(data[off + 4: off + 4 + n8], "little")
    power, ceremony_power = struct.unpack_from("<II", data, off + 4 + n8)
    return n8, modulus, power, ceremony_power


def get_powers_from_file(path: str):
    """-> (g1 (n,16) uint32 Montgomery affine, g2 (m,32) uint32 Montgomery affine)   [src/kzg/ptau.rs:347-358]"""
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        raise SetupFileError(f"FileError({e})")
    sections = parse_sections(data)
    n8, modulus, power, _ = read_header(data, sections)
    if n8 != 32 or modulus != FQ_MODULUS:
        raise SetupFileError("InvalidFieldModulus")
    if power > MAX_POWER:
        raise SetupFileError(f"ParseError(power {power} > {MAX_POWER})")
    n_g1, n_g2 = 2 * (1 << power) - 1, 1 << power
    o1, l1 = _section(sections, 2)
    o2, l2 = _section(sections, 3)
    # src/kzg/ptau.rs:251-256, 299-304: the section must hold exactly the expected number of elements
    if l1 != n_g1 * 64:
        raise SetupFileError(f"ElementSizeMismatch({n_g1 * 64}, {l1})")
    if l2 != n_g2 * 128:
        raise SetupFileError(f"ElementSizeMismatch({n_g2 * 128}, {l2})")
    g1 = np.frombuffer(data, dtype=np.uint32, count=n_g1 * 16, offset=o1).reshape(n_g1, 16).copy()
    g2 = np.frombuffer(data, dtype=np.uint32, count=n_g2 * 32, offset=o2).reshape(n_g2, 32).copy()
    return g1, g2


def write_ptau(path: str, g1_xy: np.ndarray, g2_xy: np.ndarray, power: int, ceremony_power: int = 28):
    """Writer for synthetic SRS files (so `new_from_file` can be exercised at large sizes).  Only the
    sections the reference reads carry data; the others are written empty."""
    assert g1_xy.shape == (2 * (1 << power) - 1, 16) and g2_xy.shape == (1 << power, 32)
    header = struct.pack("<I", 32) + FQ_MODULUS.to_bytes(32, "little") + struct.pack("<II", power, ceremony_power)
    payload = {1: header, 2: np.ascontiguousarray(g1_xy, np.uint32).tobytes(), 3: np.ascontiguousarray(g2_xy, np.uint32).tobytes()}
    with open(path, "wb") as f:
        f.write(FILE_TYPE + struct.pack("<II", 1, N_SECTIONS))
        for sid in SECTION_IDS:
            body = payload.get(sid, b"")
            f.write(struct.pack("<IQ", sid, len(body)) + body)
