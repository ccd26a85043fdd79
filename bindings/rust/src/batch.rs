//! Batch siblings of eccoxide's per-element API, executed by libeccbatch on B200.
//!
//! Each function is element-wise identical to the existing per-element function it names (same
//! bytes out for the same bytes in); conversions use the crate's own `to_bytes` / `from_bytes`.
//! Intended location in the reference crate: `src/batch/mod.rs` (feature `gpu-batch`).
use super::ffi::*;
use crate::curve::curve25519::{FieldElement as Fe25519, Point as EdPoint, Scalar as EdScalar};
use crate::curve::bls12_381::g1;
use crate::curve::bls12_381::scalar::Scalar as BlsScalar;
use crate::curve::field::Sign;
use crate::curve::sec2::p256r1;
use crate::protocol::ecdsa;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_long};
use std::ptr;

#[derive(Debug)]
pub enum BatchError {
    Cuda(String),
    InvalidArgument(String),
    /// `Scalar::from_bytes` would have returned `None` for this element
    NonCanonicalScalar(usize),
    /// `from_coordinate` would have returned `None` for this element
    PointNotOnCurve(usize),
    OutOfMemory,
}

/// One context per process (or per device set); thread-safe for concurrent calls.
pub struct BatchContext {
    ctx: *mut ecb_ctx,
}
unsafe impl Send for BatchContext {}
unsafe impl Sync for BatchContext {}

impl BatchContext {
    /// `devices = &[]` uses CUDA device 0.  Fails (never falls back to the CPU) without a GPU.
    pub fn new(devices: &[i32]) -> Result<Self, BatchError> {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { ecb_init(if devices.is_empty() { ptr::null() } else { devices.as_ptr() }, devices.len() as i32, &mut ctx) };
        if rc != ECB_OK {
            return Err(BatchError::Cuda("ecb_init failed: no CUDA device".into()));
        }
        Ok(BatchContext { ctx })
    }

    fn check(&self, rc: i32, bad: usize) -> Result<(), BatchError> {
        match rc {
            ECB_OK => Ok(()),
            ECB_ERR_NONCANONICAL_SCALAR => Err(BatchError::NonCanonicalScalar(bad)),
            ECB_ERR_POINT_NOT_ON_CURVE => Err(BatchError::PointNotOnCurve(bad)),
            ECB_ERR_OOM => Err(BatchError::OutOfMemory),
            ECB_ERR_INVALID_ARG => Err(BatchError::InvalidArgument(self.last_error())),
            _ => Err(BatchError::Cuda(self.last_error())),
        }
    }
    fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(ecb_last_error(self.ctx)) }.to_string_lossy().into_owned()
    }

    /// Batch sibling of `curve25519::Point::mul_base` (src/curve/curve25519.rs:840).
    pub fn ed25519_mul_base_batch(&self, scalars: &[EdScalar]) -> Result<Vec<EdPoint>, BatchError> {
        let n = scalars.len();
        let mut k = Vec::with_capacity(32 * n);
        for s in scalars {
            k.extend_from_slice(&s.to_bytes_le());
        }
        let mut xy = vec![0u8; 64 * n];
        let mut bad = usize::MAX;
        let rc = unsafe { ecb_ed25519_mul_base(self.ctx, k.as_ptr(), n, xy.as_mut_ptr(), &mut bad) };
        self.check(rc, bad)?;
        Ok(xy
            .chunks_exact(64)
            .map(|c| {
                let x = Fe25519::from_bytes_le(c[..32].try_into().unwrap()).expect("canonical");
                let y = Fe25519::from_bytes_le(c[32..].try_into().unwrap()).expect("canonical");
                EdPoint::from_coordinate(&x, &y).expect("on curve")
            })
            .collect())
    }

    /// Batch sibling of `protocol::x25519::x25519` (src/protocol/x25519.rs:36).
    pub fn x25519_batch(&self, scalars: &[[u8; 32]], us: &[[u8; 32]]) -> Result<Vec<[u8; 32]>, BatchError> {
        if scalars.len() != us.len() {
            return Err(BatchError::InvalidArgument("length mismatch".into()));
        }
        let n = scalars.len();
        let mut out = vec![[0u8; 32]; n];
        let rc = unsafe { ecb_x25519(self.ctx, scalars.as_ptr() as *const u8, us.as_ptr() as *const u8, n, out.as_mut_ptr() as *mut u8) };
        self.check(rc, usize::MAX)?;
        Ok(out)
    }

    /// Batch sibling of `&p256r1::Point * &p256r1::Scalar` (src/curve/fiat/curve_macros.rs:321).
    /// `None` entries are the point at infinity (`to_affine() == None`).
    pub fn p256r1_mul_batch(&self, points: &[p256r1::PointAffine], scalars: &[p256r1::Scalar]) -> Result<Vec<Option<p256r1::PointAffine>>, BatchError> {
        if points.len() != scalars.len() {
            return Err(BatchError::InvalidArgument("length mismatch".into()));
        }
        let n = points.len();
        let (mut k, mut xy) = (Vec::with_capacity(32 * n), Vec::with_capacity(64 * n));
        for (p, s) in points.iter().zip(scalars) {
            k.extend_from_slice(&s.to_bytes());
            let (x, y) = p.to_coordinate();
            xy.extend_from_slice(&x.to_bytes());
            xy.extend_from_slice(&y.to_bytes());
        }
        let (mut out, mut inf, mut bad) = (vec![0u8; 64 * n], vec![0u8; n], usize::MAX);
        let rc = unsafe {
            ecb_wei_mul(self.ctx, ECB_CURVE_P256R1, k.as_ptr(), xy.as_ptr(), ptr::null(), n, out.as_mut_ptr(), inf.as_mut_ptr(), &mut bad)
        };
        self.check(rc, bad)?;
        Ok(out
            .chunks_exact(64)
            .zip(inf)
            .map(|(c, i)| {
                if i != 0 {
                    return None;
                }
                let x = p256r1::FieldElement::from_bytes(c[..32].try_into().unwrap())?;
                let y = p256r1::FieldElement::from_bytes(c[32..].try_into().unwrap())?;
                p256r1::PointAffine::from_coordinate(&x, &y)
            })
            .collect())
    }

    /// Batch sibling of `ed25519::SecretKey::sign` (src/protocol/ed25519.rs:112) on raw messages.
    /// Constant-time kernels on the device; `ecb_ed25519_sign_vartime` is the fast form.
    pub fn ed25519_sign_batch(&self, seeds: &[[u8; 32]], messages: &[&[u8]]) -> Result<Vec<[u8; 64]>, BatchError> {
        if seeds.len() != messages.len() {
            return Err(BatchError::InvalidArgument("length mismatch".into()));
        }
        let n = seeds.len();
        let mut off = Vec::with_capacity(n + 1);
        let mut blob = Vec::new();
        off.push(0u64);
        for m in messages {
            blob.extend_from_slice(m);
            off.push(blob.len() as u64);
        }
        blob.push(0);
        let mut sig = vec![[0u8; 64]; n];
        let rc = unsafe {
            ecb_ed25519_sign(self.ctx, seeds.as_ptr() as *const u8, ptr::null(), blob.as_ptr(), off.as_ptr(), n, sig.as_mut_ptr() as *mut u8)
        };
        self.check(rc, usize::MAX)?;
        Ok(sig)
    }

    /// Batch sibling of `ed25519::SecretKey::public_key` (src/protocol/ed25519.rs:81).
    pub fn ed25519_public_key_batch(&self, seeds: &[[u8; 32]]) -> Result<Vec<[u8; 32]>, BatchError> {
        let n = seeds.len();
        let mut out = vec![[0u8; 32]; n];
        let rc = unsafe { ecb_ed25519_public_from_seed(self.ctx, seeds.as_ptr() as *const u8, n, out.as_mut_ptr() as *mut u8) };
        self.check(rc, usize::MAX)?;
        Ok(out)
    }

    /// Batch sibling of `p256r1::PointAffine::decompress(&x, sign)` (src/curve/fiat/curve_macros.rs:221 ->
    /// src/curve/affine.rs:48).  `None` where the reference's `CtOption` is empty.
    pub fn p256r1_decompress_batch(&self, xs: &[p256r1::FieldElement], signs: &[Sign]) -> Result<Vec<Option<p256r1::PointAffine>>, BatchError> {
        if xs.len() != signs.len() {
            return Err(BatchError::InvalidArgument("length mismatch".into()));
        }
        let n = xs.len();
        let mut xb = Vec::with_capacity(32 * n);
        for x in xs {
            xb.extend_from_slice(&x.to_bytes());
        }
        let sg: Vec<u8> = signs.iter().map(|s| matches!(s, Sign::Negative) as u8).collect();
        let (mut out, mut ok) = (vec![0u8; 64 * n], vec![0u8; n]);
        let rc = unsafe { ecb_wei_decompress(self.ctx, ECB_CURVE_P256R1, xb.as_ptr(), sg.as_ptr(), n, out.as_mut_ptr(), ok.as_mut_ptr()) };
        self.check(rc, usize::MAX)?;
        Ok(out
            .chunks_exact(64)
            .zip(ok)
            .map(|(c, present)| {
                if present == 0 {
                    return None;
                }
                let x = p256r1::FieldElement::from_bytes(c[..32].try_into().unwrap())?;
                let y = p256r1::FieldElement::from_bytes(c[32..].try_into().unwrap())?;
                p256r1::PointAffine::from_coordinate(&x, &y)
            })
            .collect())
    }

    /// Batch sibling of `bls12_381::g1::PointAffine::from_compressed` (src/curve/bls12_381/serialize.rs:286;
    /// `check_subgroup = false`: `from_compressed_oncurve_only`, :310).  The device has already validated
    /// flags, canonicity, the curve equation and (optionally) subgroup membership, so the accepted
    /// coordinates are rebuilt with the unchecked constructor path (`from_uncompressed_oncurve_only`).
    pub fn g1_from_compressed_batch(&self, encodings: &[[u8; 48]], check_subgroup: bool) -> Result<Vec<Option<g1::PointAffine>>, BatchError> {
        let n = encodings.len();
        let (mut out, mut ok) = (vec![0u8; 96 * n], vec![0u8; n]);
        let rc = unsafe {
            ecb_bls12_381_g1_from_compressed(self.ctx, encodings.as_ptr() as *const u8, n, check_subgroup as c_int, out.as_mut_ptr(), ok.as_mut_ptr())
        };
        self.check(rc, usize::MAX)?;
        Ok(out
            .chunks_exact(96)
            .zip(ok)
            .map(|(c, present)| if present == 0 { None } else { g1::PointAffine::from_uncompressed_oncurve_only(c.try_into().unwrap()) })
            .collect())
    }

    /// Batch sibling of `&g1::Point * &Scalar` on BLS12-381 G1 (src/curve/fiat/curve_macros.rs:321 ->
    /// projective.rs:842).  `in_subgroup = false` is `Point::mul` as the crate defines it: any point of E(Fp).
    /// `in_subgroup = true` is for points the caller KNOWS to be in the prime-order subgroup — everything that came
    /// out of `g1_from_compressed_batch(.., true)` or of `mul_base` — and lets the device split the scalar over the
    /// endomorphism of the crate's own subgroup test (g1.rs:105): the same group element for those points, 1.35x
    /// faster; for a point outside G1 the result would NOT be `k * P`.  The option is scoped to this call.
    pub fn g1_mul_batch(&self, points: &[g1::PointAffine], scalars: &[BlsScalar], in_subgroup: bool) -> Result<Vec<Option<g1::PointAffine>>, BatchError> {
        if points.len() != scalars.len() {
            return Err(BatchError::InvalidArgument("length mismatch".into()));
        }
        let n = points.len();
        let (mut k, mut xy) = (Vec::with_capacity(32 * n), Vec::with_capacity(96 * n));
        for (p, s) in points.iter().zip(scalars) {
            k.extend_from_slice(&s.to_bytes());
            xy.extend_from_slice(&p.to_uncompressed()); // x || y, 48 bytes each, big-endian (serialize.rs:269)
        }
        let (mut out, mut inf, mut bad) = (vec![0u8; 96 * n], vec![0u8; n], usize::MAX);
        let key = b"bls12_381_g1_glv\0".as_ptr() as *const c_char;
        let rc = unsafe {
            ecb_set_option(self.ctx, key, in_subgroup as c_long);
            let rc = ecb_wei_mul(self.ctx, ECB_CURVE_BLS12_381_G1, k.as_ptr(), xy.as_ptr(), ptr::null(), n, out.as_mut_ptr(), inf.as_mut_ptr(), &mut bad);
            ecb_set_option(self.ctx, key, 0);
            rc
        };
        self.check(rc, bad)?;
        Ok(out
            .chunks_exact(96)
            .zip(inf)
            .map(|(c, i)| if i != 0 { None } else { g1::PointAffine::from_uncompressed_oncurve_only(c.try_into().unwrap()) })
            .collect())
    }

    /// Batch sibling of `ecdsa::verify_hashed::<P256R1_*>` (src/protocol/ecdsa.rs:205): public keys,
    /// message scalars (`digest_to_scalar` output) and signatures, one bool per element.
    pub fn ecdsa_p256r1_verify_batch(
        &self,
        keys: &[p256r1::PointAffine],
        zs: &[p256r1::Scalar],
        sigs: &[ecdsa::Signature<p256r1::Scalar>],
    ) -> Result<Vec<bool>, BatchError> {
        let n = keys.len();
        if zs.len() != n || sigs.len() != n {
            return Err(BatchError::InvalidArgument("length mismatch".into()));
        }
        let (mut q, mut z, mut rs) = (Vec::with_capacity(64 * n), Vec::with_capacity(32 * n), Vec::with_capacity(64 * n));
        for i in 0..n {
            let (x, y) = keys[i].to_coordinate();
            q.extend_from_slice(&x.to_bytes());
            q.extend_from_slice(&y.to_bytes());
            z.extend_from_slice(&zs[i].to_bytes());
            rs.extend_from_slice(&sigs[i].to_bytes());
        }
        let (mut ok, mut bad) = (vec![0u8; n], usize::MAX);
        let rc = unsafe { ecb_ecdsa_verify_hashed(self.ctx, ECB_CURVE_P256R1, q.as_ptr(), z.as_ptr(), rs.as_ptr(), n, ok.as_mut_ptr(), &mut bad) };
        self.check(rc, bad)?;
        Ok(ok.into_iter().map(|b| b != 0).collect())
    }
}

impl Drop for BatchContext {
    fn drop(&mut self) {
        unsafe { ecb_destroy(self.ctx) }
    }
}
