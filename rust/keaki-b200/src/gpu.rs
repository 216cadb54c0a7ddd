//! FFI to libkeaki_b200.so (include/keaki_b200.h) and the marshalling between arkworks values and the ABI's limb
//! arrays.  arkworks keeps `Fp256<MontBackend<_, 4>>` as 4 x u64 Montgomery limbs with R = 2^256 - exactly the ABI's
//! 8 x u32 - so field elements cross without conversion; `Affine { x, y, infinity }` is not `repr(C)` and is repacked.

use ark_bn254::{Fq, Fq2, Fr, G1Affine, G1Projective, G2Affine, G2Projective};
use ark_ff::BigInt;
use std::os::raw::c_char;

#[repr(C)]
pub struct KbCtx {
    _p: [u8; 0],
}

pub const KB_OK: i32 = 0;
pub const KB_ERR_POLY_TOO_LARGE: i32 = -3;
pub const KB_ERR_INVALID_POINT: i32 = -6;

extern "C" {
    pub fn kb_ctx_create(device: i32, out: *mut *mut KbCtx) -> i32;
    pub fn kb_ctx_create_multi(devices: *const i32, ndev: i32, out: *mut *mut KbCtx) -> i32;
    pub fn kb_ctx_destroy(ctx: *mut KbCtx);
    pub fn kb_last_error(ctx: *const KbCtx) -> *const c_char;
    pub fn kb_srs_upload(ctx: *mut KbCtx, g1_xy: *const u64, n: u64, tau_g2_xy: *const u64) -> i32;
    pub fn kb_srs_validate(ctx: *mut KbCtx, first_bad: *mut u64) -> i32;
    pub fn kb_msm_g1(ctx: *mut KbCtx, scalars: *const u64, first: u64, n: u64, out_xy: *mut u64, out_inf: *mut u8) -> i32;
    pub fn kb_open_batch(ctx: *mut KbCtx, coeffs: *const u64, d: u64, points: *const u64, m: u64, proofs_xy: *mut u64, inf: *mut u8) -> i32;
    pub fn kb_open_all_fk(ctx: *mut KbCtx, coeffs: *const u64, d: u64, proofs_xy: *mut u64, inf: *mut u8) -> i32;
    pub fn kb_fr_ntt(ctx: *mut KbCtx, data: *mut u64, n: u64, inverse: i32) -> i32;
    pub fn kb_verify_batch(ctx: *mut KbCtx, com_xy: *const u64, com_inf: *const u8, points: *const u64, values: *const u64,
                           proofs_xy: *const u64, proofs_inf: *const u8, n: u64, ok: *mut u8) -> i32;
    pub fn kb_encrypt_batch(ctx: *mut KbCtx, com_xy: *const u64, com_inf: u8, points: *const u64, values: *const u64, r: *const u64,
                            msgs: *const u8, msg_off: *const u64, n: u64, ct_xy: *mut u64, ct_inf: *mut u8, msg_ct: *mut u8) -> i32;
    pub fn kb_decrypt_batch(ctx: *mut KbCtx, proofs_xy: *const u64, proofs_inf: *const u8, ct_xy: *const u64, ct_inf: *const u8,
                            msg_ct: *const u8, msg_off: *const u64, n: u64, msgs_out: *mut u8) -> i32;
}

/// Owning handle; `KZGSetup` keeps one (src/kzg.rs:22-29 has no device state: this is the only new field).
pub struct Ctx(pub *mut KbCtx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {} // the library serialises nothing: callers hold `&KZGSetup` and call one entry point at a time
impl Ctx {
    /// One GPU, or all `devices` of the box behind one handle (commit splits by point range, batches by index).
    pub fn new(devices: &[i32]) -> Self {
        let mut p = std::ptr::null_mut();
        let rc = unsafe {
            if devices.len() == 1 { kb_ctx_create(devices[0], &mut p) } else { kb_ctx_create_multi(devices.as_ptr(), devices.len() as i32, &mut p) }
        };
        assert!(rc == KB_OK && !p.is_null(), "keaki-b200: no usable sm_100 CUDA device (there is no CPU fallback)");
        Ctx(p)
    }
    pub fn check(&self, rc: i32, what: &str) {
        if rc != KB_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(kb_last_error(self.0)) }.to_string_lossy().into_owned();
            panic!("{what}: {msg} ({rc})"); // the reference unwraps / panics in the same places
        }
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { kb_ctx_destroy(self.0) }
    }
}

#[inline]
pub fn fr_limbs(x: &Fr) -> [u64; 4] {
    x.0 .0 // Montgomery limbs as held in RAM
}
#[inline]
fn fq_limbs(x: &Fq) -> [u64; 4] {
    x.0 .0
}
#[inline]
fn fq_from(l: &[u64]) -> Fq {
    Fq::new_unchecked(BigInt([l[0], l[1], l[2], l[3]])) // takes the Montgomery representation as-is (ark-ff 0.4)
}
pub fn frs(v: &[Fr]) -> Vec<u64> {
    v.iter().flat_map(fr_limbs).collect()
}
pub fn pack_g1(p: &G1Affine) -> ([u64; 8], u8) {
    if p.infinity {
        return ([0; 8], 1);
    }
    let (x, y) = (fq_limbs(&p.x), fq_limbs(&p.y));
    ([x[0], x[1], x[2], x[3], y[0], y[1], y[2], y[3]], 0)
}
pub fn unpack_g1(xy: &[u64], inf: u8) -> G1Projective {
    if inf != 0 {
        return G1Projective::default();
    }
    G1Affine::new_unchecked(fq_from(&xy[0..4]), fq_from(&xy[4..8])).into()
}
pub fn pack_g2(p: &G2Affine) -> ([u64; 16], u8) {
    let mut o = [0u64; 16];
    if p.infinity {
        return (o, 1);
    }
    for (k, c) in [&p.x.c0, &p.x.c1, &p.y.c0, &p.y.c1].iter().enumerate() {
        o[4 * k..4 * k + 4].copy_from_slice(&fq_limbs(c));
    }
    (o, 0)
}
pub fn unpack_g2(xy: &[u64], inf: u8) -> G2Projective {
    if inf != 0 {
        return G2Projective::default();
    }
    let f2 = |o: usize| Fq2::new(fq_from(&xy[o..o + 4]), fq_from(&xy[o + 4..o + 8]));
    G2Affine::new_unchecked(f2(0), f2(8)).into()
}
