//! keaki with the hot path on B200: same modules as the reference crate root (src/lib.rs:5-8) - `enc`, `kem`, `kzg`,
//! `vec` - same function names, argument order (leading `rng`), return types and error behaviour, with
//! `E = ark_bn254::Bn254` (stable Rust cannot specialise a generic `E: Pairing` onto a foreign backend).
//!
//! NOT COMPILED in this repository's build image (no Rust toolchain): reviewed source.  The ABI it binds is the one the
//! Python mirror and the C++ host layer drive in tests/.
pub mod gpu;

pub type E = ark_bn254::Bn254;
pub type Fr = ark_bn254::Fr;
pub type G1 = ark_bn254::G1Projective;
pub type G2 = ark_bn254::G2Projective;

pub mod kzg {
    use super::{gpu::*, Fr, G1, G2};
    use ark_bn254::{G1Affine, G2Affine};
    use ark_ec::CurveGroup;
    use ark_poly::{univariate::DensePolynomial, EvaluationDomain, Radix2EvaluationDomain};
    use thiserror::Error;

    /// src/kzg.rs:205-209
    #[derive(Error, Debug, PartialEq)]
    pub enum KZGError {
        #[error("Polynomial too large: {0} > {1}")]
        PolynomialTooLarge(usize, usize),
    }
    /// src/kzg/ptau.rs:360-376 (the container parser itself is unchanged host code; only the variants used here)
    #[derive(Error, Debug, PartialEq)]
    pub enum SetupFileError {
        #[error("Section is uninitialized: {0}")]
        EmptySection(u8),
        #[error("IO error: {0}")]
        ParseError(String),
    }

    /// src/kzg.rs:22-29 + the device handle
    pub struct KZGSetup {
        g1_pow: Vec<G1>,
        g1_aff: Vec<G1Affine>,
        tau_g2: G2,
        pub(crate) ctx: Ctx,
    }

    impl KZGSetup {
        fn upload(g1_aff: Vec<G1Affine>, tau_g2: G2, devices: &[i32], validate: bool) -> Result<Self, SetupFileError> {
            let ctx = Ctx::new(devices);
            let flat: Vec<u64> = g1_aff.iter().flat_map(|p| pack_g1(p).0).collect();
            let (t2, _) = pack_g2(&tau_g2.into_affine());
            ctx.check(unsafe { kb_srs_upload(ctx.0, flat.as_ptr(), g1_aff.len() as u64, t2.as_ptr()) }, "kb_srs_upload");
            if validate {
                let mut bad = 0u64;
                let rc = unsafe { kb_srs_validate(ctx.0, &mut bad) };
                if rc == KB_ERR_INVALID_POINT {
                    return Err(SetupFileError::ParseError(format!("SRS element {bad} is not a point of its group")));
                }
                ctx.check(rc, "kb_srs_validate");
            }
            let g1_pow = g1_aff.iter().map(|&p| p.into()).collect();
            Ok(Self { g1_pow, g1_aff, tau_g2, ctx })
        }
        /// src/kzg.rs:33-52.  `powers` = the output of the unchanged `ptau::get_powers_from_file` - except that the
        /// coordinates must be taken as the snarkjs Montgomery limbs they are (DESIGN.md "Deliberate deviation");
        /// the points are validated on the GPU, which the reference's `_unchecked` read never does.
        pub fn new_from_powers(g1_aff: Vec<G1Affine>, g2_aff: Vec<G2Affine>) -> Result<Self, SetupFileError> {
            let tau_g2 = g2_aff.get(1).copied().ok_or(SetupFileError::EmptySection(3))?.into();
            Self::upload(g1_aff, tau_g2, &[0], true)
        }
        /// src/kzg.rs:55-70 ("Don't use this"): host arithmetic as in the reference, then upload
        pub fn setup(secret: Fr, max_d: usize) -> Self {
            use ark_ec::AffineRepr;
            use ark_ff::Field;
            use std::ops::Mul;
            let tau_g2 = G2Affine::generator().mul(secret);
            let g1_pow: Vec<G1> = (0..max_d).map(|i| G1Affine::generator().mul(secret.pow([i as u64]))).collect();
            let g1_aff = G1::normalize_batch(&g1_pow);
            Self::upload(g1_aff, tau_g2, &[0], false).unwrap()
        }
        /// the same SRS behind ONE handle over several GPUs of the box
        pub fn on_devices(&self, devices: &[i32]) -> Self {
            Self::upload(self.g1_aff.clone(), self.tau_g2, devices, false).unwrap()
        }
        pub fn g1_pow(&self) -> &[G1] { &self.g1_pow }
        pub fn g1_aff(&self) -> &[G1Affine] { &self.g1_aff }
        pub fn tau_g2(&self) -> G2 { self.tau_g2 }
    }

    /// src/kzg.rs:89-101: `VariableBaseMSM::msm_unchecked(&setup.g1_aff, p)` -> kb_msm_g1
    pub fn commit(setup: &KZGSetup, p: &DensePolynomial<Fr>) -> Result<G1, KZGError> {
        if p.coeffs.len() > setup.g1_pow.len() {
            return Err(KZGError::PolynomialTooLarge(p.coeffs.len(), setup.g1_pow.len()));
        }
        let (mut xy, mut inf) = ([0u64; 8], 0u8);
        let s = frs(&p.coeffs);
        setup.ctx.check(unsafe { kb_msm_g1(setup.ctx.0, s.as_ptr(), 0, p.coeffs.len() as u64, xy.as_mut_ptr(), &mut inf) }, "kb_msm_g1");
        Ok(unpack_g1(&xy, inf))
    }
    /// src/kzg.rs:104-124 -> kb_open_batch (m = 1)
    pub fn open(setup: &KZGSetup, p: &DensePolynomial<Fr>, point: &Fr) -> Result<G1, KZGError> {
        if p.coeffs.len() > setup.g1_pow.len() + 1 {
            return Err(KZGError::PolynomialTooLarge(p.coeffs.len() - 1, setup.g1_pow.len()));
        }
        let (mut xy, mut inf) = ([0u64; 8], 0u8);
        let (c, z) = (frs(&p.coeffs), fr_limbs(point));
        setup.ctx.check(unsafe { kb_open_batch(setup.ctx.0, c.as_ptr(), p.coeffs.len() as u64, z.as_ptr(), 1, xy.as_mut_ptr(), &mut inf) }, "kb_open_batch");
        Ok(unpack_g1(&xy, inf))
    }
    /// src/kzg.rs:127-151 -> kb_verify_batch (n = 1)
    pub fn verify(setup: &KZGSetup, commitment: G1, point: Fr, value: Fr, proof: G1) -> Result<bool, KZGError> {
        let ((c, ci), (p, pi)) = (pack_g1(&commitment.into_affine()), pack_g1(&proof.into_affine()));
        let mut ok = 0u8;
        setup.ctx.check(unsafe { kb_verify_batch(setup.ctx.0, c.as_ptr(), &ci, fr_limbs(&point).as_ptr(), fr_limbs(&value).as_ptr(), p.as_ptr(), &pi, 1, &mut ok) },
                        "kb_verify_batch");
        Ok(ok != 0)
    }
    /// src/kzg.rs:157-203 -> kb_open_all_fk; panics like the reference when d exceeds the SRS (:169) or 2^27 (:163)
    pub fn open_fk(setup: &KZGSetup, p: &[Fr], domain_d: &Radix2EvaluationDomain<Fr>) -> Result<Vec<G1>, KZGError> {
        let d = domain_d.size();
        assert!(p.len() == d && d <= setup.g1_pow.len());
        let (mut xy, mut inf) = (vec![0u64; 8 * d], vec![0u8; d]);
        setup.ctx.check(unsafe { kb_open_all_fk(setup.ctx.0, frs(p).as_ptr(), d as u64, xy.as_mut_ptr(), inf.as_mut_ptr()) }, "kb_open_all_fk");
        Ok((0..d).map(|i| unpack_g1(&xy[8 * i..8 * i + 8], inf[i])).collect())
    }
}

pub mod kem {
    use super::{enc, kzg::KZGSetup, Fr, G1, G2};
    /// src/kem.rs:13-50: one `Fr::rand(rng)` draw (:26), then the batch entry point with n = 1 and a zero message
    pub fn encapsulate(rng: &mut impl rand::Rng, kzg_setup: &KZGSetup, commitment: G1, point: Fr, value: Fr, msg_len: usize) -> (G2, Vec<u8>) {
        enc::encrypt(rng, kzg_setup, commitment, point, value, &vec![0u8; msg_len])
    }
    /// src/kem.rs:55-72
    pub fn decapsulate(kzg_setup: &KZGSetup, proof: G1, ciphertext: G2, msg_len: usize) -> Vec<u8> {
        enc::decrypt(kzg_setup, proof, &(ciphertext, vec![0u8; msg_len]))
    }
}

pub mod enc {
    use super::{kzg::KZGSetup, vec, Fr, G1, G2};
    /// src/enc.rs:13
    pub type Ciphertext = (G2, Vec<u8>);
    /// src/enc.rs:19-40
    pub fn encrypt(rng: &mut impl rand::Rng, kzg_setup: &KZGSetup, com: G1, point: Fr, value: Fr, msg: &[u8]) -> Ciphertext {
        vec::vec_encrypt(rng, kzg_setup, com, &[point], &[value], &[msg]).pop().unwrap()
    }
    /// src/enc.rs:44-55 (the reference needs no setup here; the handle that owns the GPU is passed instead)
    pub fn decrypt(kzg_setup: &KZGSetup, proof: G1, ct: &Ciphertext) -> Vec<u8> {
        vec::vec_decrypt(kzg_setup, &[proof], &[ct]).pop().unwrap()
    }
}

pub mod vec {
    use super::{enc::Ciphertext, gpu::*, kzg::{commit, open_fk, KZGSetup}, Fr, G1};
    use ark_ec::CurveGroup;
    use ark_poly::{univariate::DensePolynomial, DenseUVPolynomial, EvaluationDomain, Radix2EvaluationDomain};
    use ark_std::UniformRand;

    /// src/vec.rs:18
    pub const PADDING_LEN: usize = 1;

    /// src/vec.rs:22-49: pad with one `Fr::rand` (:32), iFFT on the GPU, open_fk, commit
    pub fn vec_commit(rng: &mut impl rand::Rng, kzg_setup: &KZGSetup, vec: &[Fr]) -> Result<(G1, Vec<G1>), &'static str> {
        let d = vec.len() + PADDING_LEN;
        let mut padded: Vec<Fr> = vec.to_vec();
        for _ in 0..PADDING_LEN {
            padded.push(Fr::rand(rng));
        }
        let domain = Radix2EvaluationDomain::<Fr>::new(d).unwrap();
        padded.resize(domain.size(), Fr::from(0u64));
        let mut limbs = frs(&padded);
        kzg_setup.ctx.check(unsafe { kb_fr_ntt(kzg_setup.ctx.0, limbs.as_mut_ptr(), domain.size() as u64, 1) }, "kb_fr_ntt");
        let p_coeff: Vec<Fr> = limbs.chunks(4).map(|l| Fr::new_unchecked(ark_ff::BigInt([l[0], l[1], l[2], l[3]]))).collect();
        let proofs = open_fk(kzg_setup, &p_coeff, &domain).unwrap();
        let com = commit(kzg_setup, &DensePolynomial::from_coefficients_vec(p_coeff)).unwrap();
        Ok((com, proofs))
    }

    /// src/vec.rs:52-69: the r_i are drawn in index order, one per message, exactly as the loop at :63-66 consumes the
    /// rng (-> src/kem.rs:26); the loop body is ONE kb_encrypt_batch call
    pub fn vec_encrypt(rng: &mut impl rand::Rng, kzg_setup: &KZGSetup, com: G1, points: &[Fr], values: &[Fr], messages: &[&[u8]]) -> Vec<Ciphertext> {
        let n = messages.len();
        let r: Vec<Fr> = (0..n).map(|_| Fr::rand(rng)).collect();
        let (pts, vals) = (frs(&points[..n]), frs(&values[..n])); // out-of-range indexing panics where the reference's does (:64)
        let mut off = vec![0u64; n + 1];
        for i in 0..n {
            off[i + 1] = off[i] + messages[i].len() as u64;
        }
        let flat: Vec<u8> = messages.concat();
        let (c, ci) = pack_g1(&com.into_affine());
        let (mut ct, mut ct_inf, mut msg_ct) = (vec![0u64; 16 * n], vec![0u8; n], vec![0u8; flat.len().max(1)]);
        kzg_setup.ctx.check(unsafe {
            kb_encrypt_batch(kzg_setup.ctx.0, c.as_ptr(), ci, pts.as_ptr(), vals.as_ptr(), frs(&r).as_ptr(), flat.as_ptr(), off.as_ptr(), n as u64,
                             ct.as_mut_ptr(), ct_inf.as_mut_ptr(), msg_ct.as_mut_ptr())
        }, "kb_encrypt_batch");
        (0..n).map(|i| (unpack_g2(&ct[16 * i..16 * i + 16], ct_inf[i]), msg_ct[off[i] as usize..off[i + 1] as usize].to_vec())).collect()
    }

    /// src/vec.rs:72-81 -> ONE kb_decrypt_batch call
    pub fn vec_decrypt(kzg_setup: &KZGSetup, proofs: &[G1], cts: &[&Ciphertext]) -> Vec<Vec<u8>> {
        let n = cts.len();
        let (mut pxy, mut pinf, mut cxy, mut cinf) = (Vec::with_capacity(8 * n), vec![0u8; n], Vec::with_capacity(16 * n), vec![0u8; n]);
        let mut off = vec![0u64; n + 1];
        for i in 0..n {
            let (p, pi) = pack_g1(&proofs[i].into_affine());
            let (c, ci) = pack_g2(&cts[i].0.into_affine());
            pxy.extend_from_slice(&p); pinf[i] = pi; cxy.extend_from_slice(&c); cinf[i] = ci;
            off[i + 1] = off[i] + cts[i].1.len() as u64;
        }
        let flat: Vec<u8> = cts.iter().flat_map(|c| c.1.iter().copied()).collect();
        let mut out = vec![0u8; flat.len().max(1)];
        kzg_setup.ctx.check(unsafe {
            kb_decrypt_batch(kzg_setup.ctx.0, pxy.as_ptr(), pinf.as_ptr(), cxy.as_ptr(), cinf.as_ptr(), flat.as_ptr(), off.as_ptr(), n as u64, out.as_mut_ptr())
        }, "kb_decrypt_batch");
        (0..n).map(|i| out[off[i] as usize..off[i + 1] as usize].to_vec()).collect()
    }
}
