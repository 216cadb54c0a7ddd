#!/usr/bin/env python3
"""kb_encrypt_batch latency against batch size: one warp per message (pairing_warp.cu) vs one thread per message (we.cu).
Device ms of the call (warm commitment)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from keaki_b200 import _ffi  # noqa: E402

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
rng = np.random.default_rng(1)


def rand_fr(n):
    out = np.zeros((n, 8), np.uint32)
    for i in range(n):
        x = int.from_bytes(rng.bytes(40), "little") % R
        for k in range(8):
            out[i, k] = (x >> (32 * k)) & 0xFFFFFFFF
    return out


def main():
    sizes = [1, 64, 512, 1024, 2048, 4096, 8192]
    os.environ["KB_ENCRYPT_WARP_MAX"] = str(1 << 20)
    cw = _ffi.Context(0)
    os.environ["KB_ENCRYPT_WARP_MAX"] = "0"
    ct = _ffi.Context(0)
    nmax = max(sizes)
    one = np.zeros(8, np.uint32); one[0] = 7
    for c in (cw, ct):
        c.srs_generate(one, 64, download=False)
    off = np.arange(nmax + 1, dtype=np.uint64) * 32
    com, _ = cw.g1_mul_gen_batch(rand_fr(1))
    pts, rs = rand_fr(nmax), rand_fr(nmax)
    for name, vals in (("general values", rand_fr(nmax)), ("bit values", np.zeros((nmax, 8), np.uint32))):
        print(name, "\nn      warp_ms  thread_ms")
        for n in sizes:
            row = []
            for c in (cw, ct):
                best = 1e9
                for _ in range(3):
                    c.encrypt_batch(com[0], 0, pts[:n], vals[:n], rs[:n], np.zeros(32 * n, np.uint8), off[: n + 1])
                    best = min(best, c.last_kernel_ms(3))
                row.append(best)
            print("%-6d %8.3f %9.3f" % (n, row[0], row[1]), flush=True)


if __name__ == "__main__":
    main()
