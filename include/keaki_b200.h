/* keaki_b200 — C ABI of the B200-native hot path of brech1/keaki.
 *
 * The reference has no FFI; its boundary to the arithmetic is static trait dispatch into arkworks
 * (SURVEY.md §8b).  Each entry point below replaces the arkworks call(s) behind one keaki function
 * and is what a `build.rs`-linked Rust shim binds (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - Everything little-endian.  Field elements are 8 x u32 limbs in MONTGOMERY form with R = 2^256:
 *    exactly the bytes arkworks keeps in RAM for `Fp256<MontBackend<_,4>>` (4 x u64), so the shim
 *    passes `fe.0.0` without conversion.
 *  - G1 affine = x||y (16 limbs); G2 affine = x.c0||x.c1||y.c0||y.c1 (32 limbs).  Infinity is carried
 *    in a separate u8 flag array (1 = infinity, coordinates ignored / written as zero).
 *  - Pointers may be host memory (pageable or pinned) or device memory of the context's GPU (e.g. a
 *    torch tensor's data_ptr); device buffers skip the PCIe copies.  Outputs follow the same rule.
 *  - Every call is blocking and returns 0 on success or a negative kb_status; nothing unwinds.
 *    `kb_last_error` gives a message.  One context = one caller thread at a time; a context is one
 *    GPU (kb_ctx_create) or several GPUs of one box (kb_ctx_create_multi); distinct contexts are
 *    independent (one process per GPU under torchrun, or several contexts in one process).
 *  - There is NO CPU fallback: if no CUDA device is usable, kb_ctx_create fails.
 */
#ifndef KEAKI_B200_H
#define KEAKI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kb_ctx kb_ctx;

typedef enum {
  KB_OK = 0,
  KB_ERR_CUDA = -1,            /* CUDA runtime error; see kb_last_error */
  KB_ERR_ARG = -2,             /* null pointer / bad size */
  KB_ERR_POLY_TOO_LARGE = -3,  /* KZGError::PolynomialTooLarge (src/kzg.rs:93-95, 205-209) */
  KB_ERR_NO_SRS = -4,          /* kb_srs_upload / kb_srs_generate not called yet */
  KB_ERR_DOMAIN = -5,          /* size not a power of two, or > 2^28 (src/kzg.rs:163 unwrap) */
  KB_ERR_INVALID_POINT = -6    /* kb_srs_validate: an SRS element is not a point of its group */
} kb_status;

/* Library version and the CUDA architecture the kernels were built for ("sm_100a"). */
const char* kb_version(void);

/* Context on CUDA device `device`.  Replaces nothing in the reference (it has no device state);
 * owned by the shim's `KZGSetup` (src/kzg.rs:22-29). */
int32_t kb_ctx_create(int32_t device, kb_ctx** out);
/* One context over ndev GPUs of the box ("one process drives all 8 GPUs"), one host thread + stream per
 * device.  Every device holds the SRS and its tables; kb_msm_g1 (`commit`, src/kzg.rs:89-101) then splits its call by
 * point range - one partial sum per device, added on the first - and kb_encrypt_batch / kb_decrypt_batch (the loops of
 * src/vec.rs:63-66,75-78) by index; kb_srs_upload / kb_srs_generate reach every device; all other entry points run on
 * devices[0].  Same handle type and the same calls as a single-device context. */
int32_t kb_ctx_create_multi(const int32_t* devices, int32_t ndev, kb_ctx** out);
int32_t kb_ctx_device_count(const kb_ctx* ctx);
void kb_ctx_destroy(kb_ctx* ctx);
const char* kb_last_error(const kb_ctx* ctx);

/* SRS upload: `KZGSetup::setup` / `new_from_file` results (src/kzg.rs:33-70): n affine G1 powers
 * [tau^i]_1 and [tau]_2.  Builds the fixed-base tables for G2 / tau_2 used by kb_encrypt_batch. */
int32_t kb_srs_upload(kb_ctx* ctx, const uint32_t* g1_aff_xy /* n*16 */, uint64_t n,
                      const uint32_t* tau_g2_xy /* 32 */);

/* Synthetic SRS generated on the device from a known secret (harness + `KZGSetup::setup`,
 * src/kzg.rs:55-70): g1[i] = tau^(first_power + i) * G1 for i < n, tau_2 = tau * G2.  first_power = 0
 * is the reference semantics; a rank holding the point range [k*n, (k+1)*n) of a larger SRS passes
 * first_power = k*n.  If out_g1_xy / out_tau_g2_xy are non-null the points are
 * also written there. */
int32_t kb_srs_generate(kb_ctx* ctx, const uint32_t* tau /* 8, Montgomery Fr */, uint64_t first_power, uint64_t n,
                        uint32_t* out_g1_xy /* n*16 or NULL */, uint32_t* out_tau_g2_xy /* 32 or NULL */);

uint64_t kb_srs_len(const kb_ctx* ctx);

/* Validation of the resident SRS on the GPU.  The reference reads the ptau sections with
 * `deserialize_uncompressed_unchecked` (src/kzg/ptau.rs:266,314) and so accepts anything; a shim's `new_from_file` calls
 * this after kb_srs_upload.  Every G1 power must be a finite point of y^2 = x^3 + 3 with coordinates below q (cofactor 1:
 * on the curve = in the group); [tau]_2 a point of the twist with [r]P = O.  Returns KB_OK, or KB_ERR_INVALID_POINT with
 * *first_bad (may be NULL) = the smallest failing G1 index, or kb_srs_len when only [tau]_2 fails. */
int32_t kb_srs_validate(kb_ctx* ctx, uint64_t* first_bad);

/* `commit` (src/kzg.rs:89-101) = VariableBaseMSM::msm_unchecked(&g1_aff[..n], scalars):
 * sum_{i<n} scalars[i] * g1[first + i].  `first` > 0 is used for point-range sharding across GPUs
 * (SURVEY.md §8e); the reference semantics are first = 0.
 * Returns KB_ERR_POLY_TOO_LARGE if first + n > srs length. */
int32_t kb_msm_g1(kb_ctx* ctx, const uint32_t* scalars /* n*8 */, uint64_t first, uint64_t n,
                  uint32_t out_xy[16], uint8_t* out_inf);

/* out[i] = scalars[i] * G1 generator (batched fixed-base scalar multiplication: the `G1Affine::generator().mul(..)`
 * of src/kzg.rs:57,135 and src/kem.rs:22; also the harness's trapdoor-proof generator). */
int32_t kb_g1_mul_gen_batch(kb_ctx* ctx, const uint32_t* scalars /* n*8 */, uint64_t n,
                            uint32_t* out_xy /* n*16 */, uint8_t* out_inf /* n */);

/* Sum of n affine G1 points (combining per-GPU MSM partials after the gather). */
int32_t kb_g1_sum(kb_ctx* ctx, const uint32_t* pts_xy /* n*16 */, const uint8_t* inf /* n or NULL */,
                  uint64_t n, uint32_t out_xy[16], uint8_t* out_inf);

/* `open` (src/kzg.rs:104-124) for m points of one polynomial with d coefficients: proofs[j] =
 * commit((p - p(z_j)) / (x - z_j)).  KB_ERR_POLY_TOO_LARGE if d - 1 > srs length. */
int32_t kb_open_batch(kb_ctx* ctx, const uint32_t* coeffs /* d*8 */, uint64_t d,
                      const uint32_t* points /* m*8 */, uint64_t m,
                      uint32_t* proofs_xy /* m*16 */, uint8_t* proofs_inf /* m */);

/* `open_fk` (src/kzg.rs:157-203): all d openings at the d-th roots of unity; d a power of two,
 * d <= srs length.  proofs[i] is the opening at omega_d^i. */
int32_t kb_open_all_fk(kb_ctx* ctx, const uint32_t* coeffs /* d*8 */, uint64_t d,
                       uint32_t* proofs_xy /* d*16 */, uint8_t* proofs_inf /* d */);

/* Radix-2 (i)FFT over Fr in place, natural order in and out, size n = 2^k
 * (`Radix2EvaluationDomain::{fft,ifft}`, src/vec.rs:37, src/kzg.rs:185). */
int32_t kb_fr_ntt(kb_ctx* ctx, uint32_t* data /* n*8 */, uint64_t n, int32_t inverse);

/* `vec_encrypt` / `encrypt` / `encapsulate` (src/vec.rs:52-69, src/enc.rs:19-40, src/kem.rs:13-50)
 * over n messages sharing one commitment.  r[i] is the i-th `Fr::rand` draw (src/kem.rs:26), made
 * by the caller in index order.  msg_off has n+1 entries; message i is msgs[off[i]..off[i+1]).
 * Outputs: ct_g2 (affine r_i*(tau_2 - point_i*G2)), ct_inf, msg_ct = key_i XOR msg_i. */
int32_t kb_encrypt_batch(kb_ctx* ctx, const uint32_t com_xy[16], uint8_t com_inf,
                         const uint32_t* points /* n*8 */, const uint32_t* values /* n*8 */,
                         const uint32_t* r /* n*8 */, const uint8_t* msgs, const uint64_t* msg_off /* n+1 */,
                         uint64_t n, uint32_t* ct_g2_xy /* n*32 */, uint8_t* ct_inf /* n */, uint8_t* msg_ct);

/* `vec_decrypt` / `decrypt` / `decapsulate` (src/vec.rs:72-81, src/enc.rs:44-55, src/kem.rs:55-72). */
int32_t kb_decrypt_batch(kb_ctx* ctx, const uint32_t* proofs_xy /* n*16 */, const uint8_t* proofs_inf /* n or NULL */,
                         const uint32_t* ct_g2_xy /* n*32 */, const uint8_t* ct_inf /* n or NULL */,
                         const uint8_t* msg_ct, const uint64_t* msg_off /* n+1 */, uint64_t n, uint8_t* msgs_out);

/* `E::pairing` (src/kem.rs:30,58; src/kzg.rs:148) over n pairs -> n x 384 B ark-serialize GT bytes. */
int32_t kb_pairing_batch(kb_ctx* ctx, const uint32_t* g1_xy /* n*16 */, const uint8_t* g1_inf,
                         const uint32_t* g2_xy /* n*32 */, const uint8_t* g2_inf, uint64_t n,
                         uint8_t* gt_bytes /* n*384 */);

/* `verify` (src/kzg.rs:127-151) over n (commitment, point, value, proof) tuples -> ok[i] in {0,1}. */
int32_t kb_verify_batch(kb_ctx* ctx, const uint32_t* com_xy /* n*16 */, const uint8_t* com_inf,
                        const uint32_t* points /* n*8 */, const uint32_t* values /* n*8 */,
                        const uint32_t* proofs_xy /* n*16 */, const uint8_t* proofs_inf, uint64_t n,
                        uint8_t* ok /* n */);

/* Wire format: ark-serialize 0.4.2 `CanonicalSerialize` bytes of affine points - what a
 * `Ciphertext<E>` (src/enc.rs:13, `(E::G2, Vec<u8>)`) and the opening proofs of tests/laconic_ot.rs:60-75 travel in
 * between sender and receiver (`serialize_compressed` / `serialize_uncompressed` of ark-ec 0.4.2
 * models/short_weierstrass).  compress != 0: x with the SWFlags in the two top bits of the last byte (G1 32 B,
 * G2 64 B); compress == 0: x || y (64 / 128 B).  Field elements are 32-byte little-endian canonical integers. */
int32_t kb_g1_serialize(kb_ctx* ctx, const uint32_t* xy /* n*16 */, const uint8_t* inf /* n or NULL */, uint64_t n,
                        int32_t compress, uint8_t* out);
int32_t kb_g2_serialize(kb_ctx* ctx, const uint32_t* xy /* n*32 */, const uint8_t* inf /* n or NULL */, uint64_t n,
                        int32_t compress, uint8_t* out);
/* `deserialize_compressed` / `deserialize_uncompressed` (validate != 0: curve equation and, for G2, the r-torsion
 * subgroup) or their `_unchecked` forms (validate == 0).  ok[i] = 0 where arkworks returns
 * SerializationError::InvalidData (integer not below q, both flag bits, no square root, failed validation); such
 * elements come back as infinity. */
int32_t kb_g1_deserialize(kb_ctx* ctx, const uint8_t* bytes, uint64_t n, int32_t compress, int32_t validate,
                          uint32_t* xy /* n*16 */, uint8_t* inf /* n */, uint8_t* ok /* n */);
int32_t kb_g2_deserialize(kb_ctx* ctx, const uint8_t* bytes, uint64_t n, int32_t compress, int32_t validate,
                          uint32_t* xy /* n*32 */, uint8_t* inf /* n */, uint8_t* ok /* n */);

/* Test hook: elementwise field ops on the device (field: 0 = Fq, 1 = Fr; op: 0 add, 1 sub, 2 mul,
 * 3 neg, 4 inv, 5 from_mont, 6 to_mont, 7 sqr) — lets the GPU tests check the PTX primitives
 * limb-for-limb against the oracle. */
int32_t kb_debug_fp_op(kb_ctx* ctx, int32_t field, int32_t op, const uint32_t* a, const uint32_t* b,
                       uint32_t* out, uint64_t n);

/* Number of kernels this context has launched so far (bench.py reports the delta as gpu_launches). */
uint64_t kb_launch_count(const kb_ctx* ctx);

/* Device time in ms of the last call's dominant kernel(s), measured with CUDA events on the
 * context's stream: which = 0 total of all kernels of the last call, 1 = MSM bucket accumulation,
 * 2 = pairing kernel, 3 = encrypt kernel, 4 = per-commitment setup inside kb_encrypt_batch (the
 * pairing e(com, G2), its window bases and the two power tables; -1 when the call reused the cached
 * commitment).  Negative if not recorded. */
float kb_last_kernel_ms(const kb_ctx* ctx, int32_t which);

#ifdef __cplusplus
}
#endif
#endif /* KEAKI_B200_H */
