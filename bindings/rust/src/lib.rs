//! Device-resident drop-in for `toy_heaan_ckks::rings::backends::rns_ntt::{RnsBasis, RnsPoly}`
//! (src/rings/backends/rns_ntt/{basis,poly}.rs) over the C ABI of include/ckks_b200.h.
//!
//! Same names, same trait impls (`PolyRing`, `PolySampler`, `PolyAutomorphism`, `AddAssign`, `MulAssign`,
//! `Neg`, `Clone`), same errors (`RnsNttError`), so `CkksEngine`, key generation and the encoder run
//! unchanged on top.  All randomness stays in Rust: samplers draw on the host exactly as the reference
//! does (src/math/sampling.rs) and upload the limbs.
//!
//! NOTE: written against the reference's public API but not compiled in the image this library was
//! built in (no Rust toolchain there).
#![allow(clippy::missing_safety_doc)]

use std::ops::{AddAssign, MulAssign, Neg};
use std::ptr::{self, NonNull};
use std::sync::{Arc, OnceLock};

use rand::Rng;
use rand_distr::{Distribution, Normal};
use toy_heaan_ckks::math::{ternary_coefficients, uniform_coefficients};
use toy_heaan_ckks::rings::backends::rns_ntt::RnsNttError;
use toy_heaan_ckks::rings::traits::{PolyAutomorphism, PolyRing, PolySampler};
// (PolyAutomorphism is not re-exported from `rings`, only from `rings::traits`: src/rings/mod.rs:6)

pub mod ffi {
    #[repr(C)]
    pub struct CkksCtx {
        _p: [u8; 0],
    }
    #[repr(C)]
    pub struct CkksPoly {
        _p: [u8; 0],
    }
    #[repr(C)]
    pub struct CkksKsk {
        _p: [u8; 0],
    }
    #[repr(C)]
    pub struct CkksLshard {
        _p: [u8; 0],
    }
    #[repr(C)]
    pub struct CkksComm {
        _p: [u8; 0],
    }
    #[repr(C)]
    pub struct CkksCommKsk {
        _p: [u8; 0],
    }
    extern "C" {
        pub fn ckks_status_str(status: i32) -> *const std::os::raw::c_char;
        pub fn ckks_ctx_create(n: u64, moduli: *const u64, l: usize, device: i32, out: *mut *mut CkksCtx) -> i32;
        pub fn ckks_ctx_drop_last(ctx: *mut CkksCtx, k: usize, out: *mut *mut CkksCtx) -> i32;
        pub fn ckks_ctx_destroy(ctx: *mut CkksCtx) -> i32;
        pub fn ckks_ctx_sync(ctx: *mut CkksCtx) -> i32;
        pub fn ckks_ctx_total_bits(ctx: *const CkksCtx) -> u32;
        pub fn ckks_ctx_reconstruct_centered_coeff(ctx: *const CkksCtx, residues: *const u64, out: *mut i64) -> i32;
        pub fn ckks_poly_alloc(ctx: *mut CkksCtx, batch: usize, out: *mut *mut CkksPoly) -> i32;
        pub fn ckks_poly_from_coeffs(ctx: *mut CkksCtx, batch: usize, c: *const i64, len: usize, out: *mut *mut CkksPoly) -> i32;
        pub fn ckks_poly_from_channels(ctx: *mut CkksCtx, batch: usize, ch: *const u64, nch: usize, ntt: i32, out: *mut *mut CkksPoly) -> i32;
        pub fn ckks_poly_download(p: *mut CkksPoly, out: *mut u64) -> i32;
        pub fn ckks_poly_clone(p: *mut CkksPoly, out: *mut *mut CkksPoly) -> i32;
        pub fn ckks_poly_free(p: *mut CkksPoly) -> i32;
        pub fn ckks_poly_is_ntt_domain(p: *const CkksPoly) -> i32;
        pub fn ckks_poly_to_ntt_domain(p: *mut CkksPoly) -> i32;
        pub fn ckks_poly_to_coeff_domain(p: *mut CkksPoly) -> i32;
        pub fn ckks_poly_add_assign(a: *mut CkksPoly, rhs: *const CkksPoly) -> i32;
        pub fn ckks_poly_mul_assign(a: *mut CkksPoly, rhs: *const CkksPoly) -> i32;
        pub fn ckks_poly_mul_assign_naive(a: *mut CkksPoly, rhs: *const CkksPoly) -> i32;
        pub fn ckks_poly_neg(a: *mut CkksPoly) -> i32;
        pub fn ckks_poly_mod_drop_last(p: *const CkksPoly, child: *mut CkksCtx, out: *mut *mut CkksPoly) -> i32;
        pub fn ckks_poly_rescale_into(p: *const CkksPoly, child: *mut CkksCtx, out: *mut *mut CkksPoly) -> i32;
        pub fn ckks_poly_automorphism(p: *const CkksPoly, e: u64, out: *mut *mut CkksPoly) -> i32;
        pub fn ckks_poly_rotate_slots(p: *const CkksPoly, k: i32, out: *mut *mut CkksPoly) -> i32;
        pub fn ckks_poly_to_coeffs(p: *const CkksPoly, out: *mut i64) -> i32;
        pub fn ckks_poly_to_coeffs_wide(p: *const CkksPoly, out_i64: *mut i64, out_f64: *mut f64, overflow: *mut i32) -> i32;
        pub fn ckks_ctx_trim(ctx: *mut CkksCtx) -> i32;
        pub fn ckks_set_nvtx(on: i32) -> i32;
        /// 0 never / 1 automatic (default) / 2 whenever possible: the gadget product through auxiliary 30-bit NTT primes
        /// (exact).  Read when a key is uploaded and when a product is computed.
        pub fn ckks_set_ks_aux(mode: i32) -> i32;
        pub fn ckks_ksk_upload(ctx: *mut CkksCtx, a: *const u64, b: *const u64, out: *mut *mut CkksKsk) -> i32;
        pub fn ckks_ksk_free(k: *mut CkksKsk) -> i32;
        pub fn ckks_ct_mul_relin(a0: *const CkksPoly, a1: *const CkksPoly, b0: *const CkksPoly, b1: *const CkksPoly,
                                 rlk: *const CkksKsk, c0: *mut *mut CkksPoly, c1: *mut *mut CkksPoly) -> i32;
        pub fn ckks_ct_rescale(c0: *const CkksPoly, c1: *const CkksPoly, child: *mut CkksCtx,
                               o0: *mut *mut CkksPoly, o1: *mut *mut CkksPoly, bits: *mut u32) -> i32;
        pub fn ckks_ct_mul_relin_rescale(a0: *const CkksPoly, a1: *const CkksPoly, b0: *const CkksPoly, b1: *const CkksPoly,
                                         rlk: *const CkksKsk, child: *mut CkksCtx, o0: *mut *mut CkksPoly, o1: *mut *mut CkksPoly) -> i32;
        pub fn ckks_ct_rotate(c0: *const CkksPoly, c1: *const CkksPoly, rotk: *const CkksKsk, k: i32,
                              o0: *mut *mut CkksPoly, o1: *mut *mut CkksPoly) -> i32;
        pub fn ckks_ctx_ntt_table(ctx: *const CkksCtx, channel: usize, which: i32, out: *mut u64) -> i32;
        // Optional limb-sharded mode (one process per GPU; INTEGRATION.md section 3b shows the call sequence).
        pub fn ckks_lshard_create(n: u64, moduli: *const u64, l: usize, rank: i32, world: i32, device: i32, chunk: usize,
                                  out: *mut *mut CkksLshard) -> i32;
        pub fn ckks_lshard_destroy(s: *mut CkksLshard) -> i32;
        pub fn ckks_lshard_drop_last(s: *mut CkksLshard, child: *mut *mut CkksLshard) -> i32;
        pub fn ckks_lshard_local_ctx(s: *mut CkksLshard) -> *mut CkksCtx;
        pub fn ckks_lshard_ipc_size() -> usize;
        pub fn ckks_lshard_ipc_export(s: *mut CkksLshard, blob: *mut std::os::raw::c_void) -> i32;
        pub fn ckks_lshard_ipc_import(s: *mut CkksLshard, blobs: *const std::os::raw::c_void) -> i32;
        pub fn ckks_lshard_ksk_upload(s: *mut CkksLshard, a: *const u64, b: *const u64, out: *mut *mut CkksKsk) -> i32;
        pub fn ckks_lshard_ct_mul_relin_rescale(s: *mut CkksLshard, a0: *const CkksPoly, a1: *const CkksPoly, b0: *const CkksPoly,
                                                b1: *const CkksPoly, rlk: *const CkksKsk, child: *mut CkksLshard,
                                                o0: *mut *mut CkksPoly, o1: *mut *mut CkksPoly) -> i32;
        pub fn ckks_lshard_ct_rotate(s: *mut CkksLshard, c0: *const CkksPoly, c1: *const CkksPoly, rotk: *const CkksKsk, k: i32,
                                     o0: *mut *mut CkksPoly, o1: *mut *mut CkksPoly) -> i32;
        pub fn ckks_lshard_check(s: *mut CkksLshard) -> i32;
        // Batch-sharded multi-GPU group: one process, a host batch cut into one contiguous share per device.
        pub fn ckks_comm_init(ndev: i32, devices: *const i32, n: u64, moduli: *const u64, l: usize, out: *mut *mut CkksComm) -> i32;
        pub fn ckks_comm_destroy(c: *mut CkksComm) -> i32;
        pub fn ckks_comm_drop_last(c: *mut CkksComm, k: usize, out: *mut *mut CkksComm) -> i32;
        pub fn ckks_comm_size(c: *const CkksComm) -> i32;
        pub fn ckks_comm_ksk_upload(c: *mut CkksComm, a: *const u64, b: *const u64, out: *mut *mut CkksCommKsk) -> i32;
        pub fn ckks_comm_ksk_free(k: *mut CkksCommKsk) -> i32;
        pub fn ckks_comm_ct_mul_relin_rescale_host(c: *mut CkksComm, rlk: *const CkksCommKsk, batch: usize, a0: *const u64, a1: *const u64,
                                                   b0: *const u64, b1: *const u64, o0: *mut u64, o1: *mut u64) -> i32;
        pub fn ckks_comm_ct_rotate_host(c: *mut CkksComm, rotk: *const CkksCommKsk, k: i32, batch: usize, c0: *const u64, c1: *const u64,
                                        o0: *mut u64, o1: *mut u64) -> i32;
        pub fn ckks_ct_mul_relin_rescale_host(ctx: *mut CkksCtx, child: *mut CkksCtx, rlk: *const CkksKsk, batch: usize,
                                              a0: *const u64, a1: *const u64, b0: *const u64, b1: *const u64,
                                              o0: *mut u64, o1: *mut u64) -> i32;
    }
}

/// Status codes 1..6 are `RnsNttError` one to one (errors.rs:3-22); everything else is a bug or a CUDA failure.
fn to_err(rc: i32, n: usize) -> RnsNttError {
    match rc {
        // variant payloads (errors.rs:5-21) that only the host knows are filled by the callers below
        1 => RnsNttError::InvalidDegree { degree: n },
        2 => RnsNttError::EmptyBasis,
        3 => RnsNttError::NonNttFriendlyModulus { modulus: 0, degree: n },
        4 => RnsNttError::InvalidModDrop { drop_count: 0, channel_count: 0 },
        5 => RnsNttError::ChannelCountMismatch { expected: 0, actual: 0 },
        6 => RnsNttError::NonReducedCoefficient { coefficient: 0, modulus: 0 },
        other => panic!("ckks_b200: {}", unsafe { std::ffi::CStr::from_ptr(ffi::ckks_status_str(other)) }.to_string_lossy()),
    }
}
fn check(rc: i32) {
    assert_eq!(rc, 0, "ckks_b200: {}", unsafe { std::ffi::CStr::from_ptr(ffi::ckks_status_str(rc)) }.to_string_lossy());
}

/// `RnsBasis<N>` (basis.rs:91-181): moduli on the host, NTT tables on the device.
pub struct RnsBasis<const N: usize> {
    ctx: NonNull<ffi::CkksCtx>,
    moduli: Vec<u64>,
}
unsafe impl<const N: usize> Send for RnsBasis<N> {}
unsafe impl<const N: usize> Sync for RnsBasis<N> {} // the handle is only read after construction
impl<const N: usize> Drop for RnsBasis<N> {
    fn drop(&mut self) {
        unsafe { ffi::ckks_ctx_destroy(self.ctx.as_ptr()) };
    }
}
/// `NttTable<N>` (basis.rs:6-17) of one channel, read back from the library in the reference's layout; `Vec`s instead
/// of the reference's `[u64; N]` stack arrays so that it also exists at N = 2^16.
pub struct NttTable<const N: usize> {
    pub modulus: u64,
    pub n_inv: u64,
    pub forward_roots: Vec<u64>,
    pub inverse_roots: Vec<u64>,
    pub twist_factors: Vec<u64>,
    pub untwist_factors: Vec<u64>,
}

/// CUDA ordinal used by `RnsBasis::new`: `CKKS_B200_DEVICE` if set, else 0.  `new_on` takes it explicitly.
pub fn default_device() -> i32 {
    std::env::var("CKKS_B200_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0)
}

impl<const N: usize> RnsBasis<N> {
    pub fn new(moduli: Vec<u64>) -> Result<Self, RnsNttError> {
        Self::new_on(moduli, default_device())
    }
    /// `RnsBasis::new` (basis.rs:97-106) with the tables built on CUDA device `device`.
    pub fn new_on(moduli: Vec<u64>, device: i32) -> Result<Self, RnsNttError> {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { ffi::ckks_ctx_create(N as u64, moduli.as_ptr(), moduli.len(), device, &mut ctx) };
        if rc == 3 {
            let bad = moduli.iter().copied().find(|&q| !toy_heaan_ckks::math::is_ntt_friendly_prime(q, N as u64)).unwrap_or(0);
            return Err(RnsNttError::NonNttFriendlyModulus { modulus: bad, degree: N });
        }
        if rc != 0 {
            return Err(to_err(rc, N));
        }
        Ok(Self { ctx: NonNull::new(ctx).unwrap(), moduli })
    }
    pub fn moduli(&self) -> &[u64] {
        &self.moduli
    }
    pub fn channel_count(&self) -> usize {
        self.moduli.len()
    }
    pub fn drop_last(&self, drop_count: usize) -> Result<Self, RnsNttError> {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { ffi::ckks_ctx_drop_last(self.ctx.as_ptr(), drop_count, &mut ctx) };
        if rc == 4 {
            return Err(RnsNttError::InvalidModDrop { drop_count, channel_count: self.moduli.len() });
        }
        if rc != 0 {
            return Err(to_err(rc, N));
        }
        let keep = self.moduli.len() - drop_count;
        Ok(Self { ctx: NonNull::new(ctx).unwrap(), moduli: self.moduli[..keep].to_vec() })
    }
    pub fn total_bits(&self) -> u32 {
        unsafe { ffi::ckks_ctx_total_bits(self.ctx.as_ptr()) }
    }
    /// `RnsBasis::ntt_table(channel)` (basis.rs:112-114).
    pub fn ntt_table(&self, channel: usize) -> NttTable<N> {
        let get = |which: i32, len: usize| -> Vec<u64> {
            let mut v = vec![0u64; len];
            check(unsafe { ffi::ckks_ctx_ntt_table(self.ctx.as_ptr(), channel, which, v.as_mut_ptr()) });
            v
        };
        NttTable {
            modulus: self.moduli[channel],
            n_inv: get(4, 1)[0],
            forward_roots: get(0, N),
            inverse_roots: get(1, N),
            twist_factors: get(2, N),
            untwist_factors: get(3, N),
        }
    }
    pub fn reconstruct_centered_coeff(&self, residues: &[u64]) -> i64 {
        let mut v = 0i64;
        check(unsafe { ffi::ckks_ctx_reconstruct_centered_coeff(self.ctx.as_ptr(), residues.as_ptr(), &mut v) });
        v
    }
}

/// `RnsPoly<N>` (poly.rs:26-30) with its limbs resident in HBM and a lazily synchronised host mirror so that
/// `channels()` keeps returning `&[[u64; N]]` (examples slice it: encrypt_mul.rs:112, horner_chain.rs:92).
pub struct RnsPoly<const N: usize> {
    h: NonNull<ffi::CkksPoly>,
    basis: Arc<RnsBasis<N>>,
    mirror: OnceLock<Vec<[u64; N]>>,
}
// The reference's RnsPoly is plain data (Vec + Arc), hence Send + Sync; so is this one: the device handle is only
// mutated through `&mut self`, `&self` methods enqueue reads on the context's stream (the CUDA stream API and the
// library's own bookkeeping are thread-safe; the host-buffer pipeline of a context tree is serialised by a mutex in
// the library), and the host mirror is a OnceLock.
unsafe impl<const N: usize> Send for RnsPoly<N> {}
unsafe impl<const N: usize> Sync for RnsPoly<N> {}
impl<const N: usize> Drop for RnsPoly<N> {
    fn drop(&mut self) {
        unsafe { ffi::ckks_poly_free(self.h.as_ptr()) };
    }
}
impl<const N: usize> Clone for RnsPoly<N> {
    fn clone(&self) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::ckks_poly_clone(self.h.as_ptr(), &mut h) });
        Self::wrap(h, self.basis.clone())
    }
}
impl<const N: usize> RnsPoly<N> {
    fn wrap(h: *mut ffi::CkksPoly, basis: Arc<RnsBasis<N>>) -> Self {
        Self { h: NonNull::new(h).unwrap(), basis, mirror: OnceLock::new() }
    }
    pub fn zero(basis: Arc<RnsBasis<N>>) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::ckks_poly_alloc(basis.ctx.as_ptr(), 1, &mut h) });
        Self::wrap(h, basis)
    }
    pub fn from_coeffs(coeffs: &[i64], basis: Arc<RnsBasis<N>>) -> Self {
        assert!(coeffs.len() >= N, "Insufficient coefficients: expected at least {}, got {}", N, coeffs.len()); // poly.rs:50-54
        let mut h = ptr::null_mut();
        check(unsafe { ffi::ckks_poly_from_coeffs(basis.ctx.as_ptr(), 1, coeffs.as_ptr(), coeffs.len(), &mut h) });
        Self::wrap(h, basis)
    }
    pub fn from_channels(channels: Vec<[u64; N]>, basis: Arc<RnsBasis<N>>, is_ntt_domain: bool) -> Result<Self, RnsNttError> {
        let mut h = ptr::null_mut();
        let rc = unsafe {
            ffi::ckks_poly_from_channels(basis.ctx.as_ptr(), 1, channels.as_ptr() as *const u64, channels.len(), is_ntt_domain as i32, &mut h)
        };
        if rc != 0 {
            return Err(to_err(rc, N));
        }
        Ok(Self::wrap(h, basis))
    }
    pub fn new_unchecked(channels: Vec<[u64; N]>, basis: Arc<RnsBasis<N>>, is_ntt_domain: bool) -> Self {
        Self::from_channels(channels, basis, is_ntt_domain).expect("new_unchecked: caller guarantees reduced channels")
    }
    pub fn channels(&self) -> &[[u64; N]] {
        self.mirror.get_or_init(|| {
            let mut out = vec![[0u64; N]; self.basis.channel_count()];
            check(unsafe { ffi::ckks_poly_download(self.h.as_ptr(), out.as_mut_ptr() as *mut u64) });
            out
        })
    }
    pub fn basis(&self) -> &Arc<RnsBasis<N>> {
        &self.basis
    }
    pub fn is_ntt_domain(&self) -> bool {
        unsafe { ffi::ckks_poly_is_ntt_domain(self.h.as_ptr()) == 1 }
    }
    pub fn to_ntt_domain(&mut self) {
        check(unsafe { ffi::ckks_poly_to_ntt_domain(self.h.as_ptr()) });
        self.mirror.take();
    }
    pub fn to_coeff_domain(&mut self) {
        check(unsafe { ffi::ckks_poly_to_coeff_domain(self.h.as_ptr()) });
        self.mirror.take();
    }
    /// Schoolbook O(N^2) product (poly.rs:339-367); both operands in the coefficient domain.
    pub fn mul_assign_naive(&mut self, rhs: &RnsPoly<N>) {
        check(unsafe { ffi::ckks_poly_mul_assign_naive(self.h.as_ptr(), rhs.h.as_ptr()) });
        self.mirror.take();
    }
    pub fn rescale_into(&self, new_basis: Arc<RnsBasis<N>>) -> Result<Self, RnsNttError> {
        let mut h = ptr::null_mut();
        let rc = unsafe { ffi::ckks_poly_rescale_into(self.h.as_ptr(), new_basis.ctx.as_ptr(), &mut h) };
        if rc != 0 {
            return Err(to_err(rc, N));
        }
        Ok(Self::wrap(h, new_basis))
    }
    /// `RnsPoly::rescale` (poly.rs:246-249): rescale_into a fresh `drop_last(1)` basis.
    pub fn rescale(&self) -> Result<Self, RnsNttError> {
        let child = Arc::new(self.basis.drop_last(1)?);
        self.rescale_into(child)
    }
    pub fn mod_drop_last(&self, drop_count: usize) -> Result<Self, RnsNttError> {
        let child = Arc::new(self.basis.drop_last(drop_count)?);
        let mut h = ptr::null_mut();
        let rc = unsafe { ffi::ckks_poly_mod_drop_last(self.h.as_ptr(), child.ctx.as_ptr(), &mut h) };
        if rc != 0 {
            return Err(to_err(rc, N));
        }
        Ok(Self::wrap(h, child))
    }
    /// Centred CRT for a basis of ANY size (the reference's `reconstruct_centered_coeff`, basis.rs:158-180, stops at
    /// Q < 2^128): the centred value of every coefficient rounded to f64, and whether some |x| >= 2^63.
    pub fn to_coeffs_wide(&self) -> (Vec<f64>, bool) {
        let mut out = vec![0f64; N];
        let mut overflow = 0i32;
        check(unsafe { ffi::ckks_poly_to_coeffs_wide(self.h.as_ptr(), ptr::null_mut(), out.as_mut_ptr(), &mut overflow) });
        (out, overflow != 0)
    }
    /// Raw handle for the batched / fused entry points (`ckks_ct_*`).
    pub fn handle(&self) -> *mut ffi::CkksPoly {
        self.h.as_ptr()
    }
}
impl<'a, const N: usize> AddAssign<&'a RnsPoly<N>> for RnsPoly<N> {
    fn add_assign(&mut self, rhs: &'a Self) {
        debug_assert!(Arc::ptr_eq(&self.basis, &rhs.basis)); // poly.rs:260-263
        check(unsafe { ffi::ckks_poly_add_assign(self.h.as_ptr(), rhs.h.as_ptr()) });
        self.mirror.take();
    }
}
impl<'a, const N: usize> MulAssign<&'a RnsPoly<N>> for RnsPoly<N> {
    fn mul_assign(&mut self, rhs: &'a Self) {
        debug_assert!(Arc::ptr_eq(&self.basis, &rhs.basis)); // poly.rs:288-291
        check(unsafe { ffi::ckks_poly_mul_assign(self.h.as_ptr(), rhs.h.as_ptr()) });
        self.mirror.take();
    }
}
impl<const N: usize> Neg for RnsPoly<N> {
    type Output = Self;
    fn neg(mut self) -> Self {
        check(unsafe { ffi::ckks_poly_neg(self.h.as_ptr()) });
        self.mirror.take();
        self
    }
}
impl<const N: usize> PolyRing<N> for RnsPoly<N> {
    type Context = Arc<RnsBasis<N>>;
    fn zero(context: &Self::Context) -> Self {
        RnsPoly::zero(context.clone())
    }
    fn from_coeffs(coeffs: &[i64], context: &Self::Context) -> Self {
        RnsPoly::from_coeffs(coeffs, context.clone())
    }
    fn to_coeffs(&self) -> [i64; N] {
        let mut out = [0i64; N];
        check(unsafe { ffi::ckks_poly_to_coeffs(self.h.as_ptr(), out.as_mut_ptr()) });
        out
    }
    fn context(&self) -> &Self::Context {
        &self.basis
    }
}
impl<const N: usize> PolySampler<N> for RnsPoly<N> {
    // Host sampling in exactly the reference's order (poly.rs:436-478), then upload.
    fn sample_uniform<R: Rng>(context: &Self::Context, rng: &mut R) -> Self {
        let mut channels = vec![[0u64; N]; context.channel_count()];
        for (ch, channel) in channels.iter_mut().enumerate() {
            *channel = uniform_coefficients::<N, _>(context.moduli()[ch], rng);
        }
        RnsPoly::new_unchecked(channels, context.clone(), false)
    }
    fn sample_gaussian<R: Rng>(std_dev: f64, context: &Self::Context, rng: &mut R) -> Self {
        let normal = Normal::new(0.0, std_dev).expect("sample_gaussian: std_dev must be finite and positive");
        let mut noise = [0i64; N];
        for n in noise.iter_mut() {
            *n = normal.sample(rng).round() as i64;
        }
        RnsPoly::from_coeffs(&noise, context.clone())
    }
    fn sample_tribits<R: Rng>(hamming_weight: usize, context: &Self::Context, rng: &mut R) -> Self {
        let ternary = ternary_coefficients::<N, _>(hamming_weight, rng);
        RnsPoly::from_coeffs(&ternary, context.clone())
    }
    fn sample_noise<R: Rng>(variance: f64, context: &Self::Context, rng: &mut R) -> Self {
        Self::sample_gaussian(variance.sqrt(), context, rng)
    }
}
impl<const N: usize> PolyAutomorphism<N> for RnsPoly<N> {
    fn automorphism(&self, exponent: u64) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::ckks_poly_automorphism(self.h.as_ptr(), exponent, &mut h) });
        Self::wrap(h, self.basis.clone())
    }
    fn rotate_slots(&self, k: i32) -> Self {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::ckks_poly_rotate_slots(self.h.as_ptr(), k, &mut h) });
        Self::wrap(h, self.basis.clone())
    }
}

/// `RnsGadgetRelinKey` / `RnsGadgetRotationKey` (engine.rs:225-253) uploaded once and kept in HBM.
pub struct DeviceGadgetKey<const N: usize> {
    k: NonNull<ffi::CkksKsk>,
    pub rotation: i32,
}
impl<const N: usize> Drop for DeviceGadgetKey<N> {
    fn drop(&mut self) {
        unsafe { ffi::ckks_ksk_free(self.k.as_ptr()) };
    }
}
impl<const N: usize> DeviceGadgetKey<N> {
    /// `a`, `b`: the key's polynomial vectors as generated by the (unchanged) engine; their limbs are gathered
    /// into [digit][limb][N] and transformed once on the device.
    pub fn upload(basis: &Arc<RnsBasis<N>>, a: &[RnsPoly<N>], b: &[RnsPoly<N>], rotation: i32) -> Self {
        let flat = |v: &[RnsPoly<N>]| -> Vec<u64> { v.iter().flat_map(|p| p.channels().iter().flatten().copied()).collect() };
        let (fa, fb) = (flat(a), flat(b));
        let mut k = ptr::null_mut();
        check(unsafe { ffi::ckks_ksk_upload(basis.ctx.as_ptr(), fa.as_ptr(), fb.as_ptr(), &mut k) });
        Self { k: NonNull::new(k).unwrap(), rotation }
    }
    pub fn handle(&self) -> *const ffi::CkksKsk {
        self.k.as_ptr()
    }
}

/// `mul_ciphertexts_gadget` + `rescale_ciphertext` (engine.rs:473-539, 263-282) in one device call.
pub fn mul_relin_rescale<const N: usize>(
    a: (&RnsPoly<N>, &RnsPoly<N>),
    b: (&RnsPoly<N>, &RnsPoly<N>),
    rlk: &DeviceGadgetKey<N>,
) -> Result<(RnsPoly<N>, RnsPoly<N>), RnsNttError> {
    let child = Arc::new(a.0.basis().drop_last(1)?);
    let (mut o0, mut o1) = (ptr::null_mut(), ptr::null_mut());
    check(unsafe { ffi::ckks_ct_mul_relin_rescale(a.0.handle(), a.1.handle(), b.0.handle(), b.1.handle(), rlk.handle(), child.ctx.as_ptr(), &mut o0, &mut o1) });
    Ok((RnsPoly::wrap(o0, child.clone()), RnsPoly::wrap(o1, child)))
}
/// `rotate_ciphertext` (engine.rs:412-463).
pub fn rotate<const N: usize>(ct: (&RnsPoly<N>, &RnsPoly<N>), rotk: &DeviceGadgetKey<N>) -> (RnsPoly<N>, RnsPoly<N>) {
    let (mut o0, mut o1) = (ptr::null_mut(), ptr::null_mut());
    check(unsafe { ffi::ckks_ct_rotate(ct.0.handle(), ct.1.handle(), rotk.handle(), rotk.rotation, &mut o0, &mut o1) });
    (RnsPoly::wrap(o0, ct.0.basis().clone()), RnsPoly::wrap(o1, ct.0.basis().clone()))
}

/// A `Vec<Ciphertext>` worth of host limbs spread over the GPUs of the box in ONE call (`ckks_comm_*`): what replaces
/// the reference's serial loop over ciphertexts (examples/horner_chain.rs:211-278).  `cts`: (c0, c1) channel vectors
/// per ciphertext, level L; returns the rescaled ciphertexts at level L-1.
pub struct BatchShard<const N: usize> {
    comm: NonNull<ffi::CkksComm>,
    channel_count: usize,
}
unsafe impl<const N: usize> Send for BatchShard<N> {}
impl<const N: usize> Drop for BatchShard<N> {
    fn drop(&mut self) {
        unsafe { ffi::ckks_comm_destroy(self.comm.as_ptr()) };
    }
}
impl<const N: usize> BatchShard<N> {
    pub fn new(moduli: &[u64], devices: &[i32]) -> Result<Self, RnsNttError> {
        let mut c = ptr::null_mut();
        let rc = unsafe { ffi::ckks_comm_init(devices.len() as i32, devices.as_ptr(), N as u64, moduli.as_ptr(), moduli.len(), &mut c) };
        if rc != 0 {
            return Err(to_err(rc, N));
        }
        Ok(Self { comm: NonNull::new(c).unwrap(), channel_count: moduli.len() })
    }
    pub fn mul_relin_rescale(
        &self,
        rlk: (&[RnsPoly<N>], &[RnsPoly<N>]),
        a: &[(Vec<[u64; N]>, Vec<[u64; N]>)],
        b: &[(Vec<[u64; N]>, Vec<[u64; N]>)],
    ) -> Result<Vec<(Vec<[u64; N]>, Vec<[u64; N]>)>, RnsNttError> {
        assert_eq!(a.len(), b.len());
        let l = self.channel_count;
        let flat_key = |v: &[RnsPoly<N>]| -> Vec<u64> { v.iter().flat_map(|p| p.channels().iter().flatten().copied()).collect() };
        let (ka, kb) = (flat_key(rlk.0), flat_key(rlk.1));
        let mut key = ptr::null_mut();
        let rc = unsafe { ffi::ckks_comm_ksk_upload(self.comm.as_ptr(), ka.as_ptr(), kb.as_ptr(), &mut key) };
        if rc != 0 {
            return Err(to_err(rc, N));
        }
        let gather = |cts: &[(Vec<[u64; N]>, Vec<[u64; N]>)], second: bool| -> Vec<u64> {
            cts.iter().flat_map(|ct| (if second { &ct.1 } else { &ct.0 }).iter().flatten().copied()).collect()
        };
        let (a0, a1, b0, b1) = (gather(a, false), gather(a, true), gather(b, false), gather(b, true));
        let out_words = a.len() * (l - 1) * N;
        let (mut o0, mut o1) = (vec![0u64; out_words], vec![0u64; out_words]);
        let rc = unsafe {
            ffi::ckks_comm_ct_mul_relin_rescale_host(self.comm.as_ptr(), key, a.len(), a0.as_ptr(), a1.as_ptr(), b0.as_ptr(), b1.as_ptr(), o0.as_mut_ptr(), o1.as_mut_ptr())
        };
        unsafe { ffi::ckks_comm_ksk_free(key) };
        if rc != 0 {
            return Err(to_err(rc, N));
        }
        let split = |flat: &[u64]| -> Vec<Vec<[u64; N]>> {
            flat.chunks((l - 1) * N).map(|ct| ct.chunks(N).map(|limb| <[u64; N]>::try_from(limb).unwrap()).collect()).collect()
        };
        Ok(split(&o0).into_iter().zip(split(&o1)).collect())
    }
}
