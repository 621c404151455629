//! Safe wrapper that keeps the reference's KZG surface
//! (myzkp/src/modules/algebra/kzg.rs:27-72, gemini.rs:112-114) and moves the work to the GPU.
//!
//! Marshalling follows the reference's only host<->device precedent,
//! myzkp/examples/sumcheck/src/utils.rs:51-72: canonical value -> to_u64_digits -> 32 bytes LE.
use std::ffi::CStr;
use std::ptr;

use myzkp::modules::algebra::curve::bn128::{Fq, Fq2, FqOrder, G1Point, G2Point, BN128};
use myzkp::modules::algebra::field::Field;
use myzkp::modules::algebra::gemini::SplitFoldError;
use myzkp::modules::algebra::kzg::PublicKeyKZG;
use myzkp::modules::algebra::polynomial::Polynomial;
use myzkp::modules::algebra::ring::Ring;
use myzkp_b200_sys as sys;
use num_bigint::{BigInt, Sign};
use num_traits::Zero;

pub type CommitmentKZG = G1Point;

pub struct ProofKZG {
    pub y: FqOrder,
    pub w: G1Point,
}

/// PublicKeyKZG whose powers_1 live on the GPU as the resident SRS table (kzg.rs:8-11).
pub struct GpuPublicKeyKZG {
    ctx: *mut sys::myzkp_ctx,
    pub powers_2: Vec<G2Point>,
}

impl Drop for GpuPublicKeyKZG {
    fn drop(&mut self) {
        unsafe { sys::myzkp_ctx_destroy(self.ctx) };
    }
}

// A context may be used from any ONE thread at a time (every entry point selects its device first), so the handle
// can move between threads - e.g. one thread per rank of the range-sharded prover.  It is not Sync.
unsafe impl Send for GpuPublicKeyKZG {}

/// The reference's public key (kzg.rs:8-11) moved to the GPU: powers_1 are uploaded once and become the resident
/// SRS table (myzkp_srs_load_g1), powers_2 stay on the host for the verifier.  This is what lets existing callers
/// that hold a `PublicKeyKZG` (das/avail.rs:38,96,132, das/eigenda.rs:44,99,119) switch to the device path.
impl From<&PublicKeyKZG> for GpuPublicKeyKZG {
    fn from(pk: &PublicKeyKZG) -> Self {
        let mut ctx = ptr::null_mut();
        let code = unsafe { sys::myzkp_ctx_create(&mut ctx, 0) };
        assert!(code == sys::MYZKP_OK, "no usable CUDA device (there is no CPU fallback)");
        let mut bytes = Vec::with_capacity(64 * pk.powers_1.len());
        for p in &pk.powers_1 {
            bytes.extend_from_slice(&g1_to_bytes(p));
        }
        check(ctx, unsafe { sys::myzkp_srs_load_g1(ctx, bytes.as_ptr(), pk.powers_1.len()) });
        GpuPublicKeyKZG { ctx, powers_2: pk.powers_2.clone() }
    }
}

impl GpuPublicKeyKZG {
    /// powers_1 read back from the device (kzg.rs:9)
    pub fn powers_1(&self) -> Vec<G1Point> {
        let n = unsafe { sys::myzkp_srs_len(self.ctx) };
        let mut out = vec![0u8; 64 * n];
        check(self.ctx, unsafe { sys::myzkp_srs_read_g1(self.ctx, 0, n, out.as_mut_ptr()) });
        out.chunks_exact(64).map(|c| point_from_bytes(c.try_into().unwrap())).collect()
    }
    /// the same key in the reference's own type (e.g. to hand to the reference's verifier)
    pub fn to_reference(&self) -> PublicKeyKZG {
        PublicKeyKZG { powers_1: self.powers_1(), powers_2: self.powers_2.clone() }
    }
}

fn check(ctx: *mut sys::myzkp_ctx, code: i32) {
    if code != sys::MYZKP_OK {
        let msg = unsafe { CStr::from_ptr(sys::myzkp_last_error(ctx)) }.to_string_lossy().into_owned();
        // the reference panics in the corresponding situations (e.g. polynomial.rs:162)
        panic!("myzkp_b200 error {}: {}", code, msg);
    }
}

fn scalar_to_le(x: &FqOrder) -> [u8; 32] {
    let (_, digits) = x.sanitize().get_value().to_u64_digits(); // polynomial.rs:162 sanitizes before the MSM
    let mut out = [0u8; 32];
    for (i, d) in digits.iter().take(4).enumerate() {
        out[8 * i..8 * i + 8].copy_from_slice(&d.to_le_bytes());
    }
    out
}

fn marshal_scalars(coef: &[FqOrder]) -> Vec<u8> {
    let mut v = Vec::with_capacity(32 * coef.len());
    for c in coef {
        v.extend_from_slice(&scalar_to_le(c));
    }
    v
}

fn point_from_bytes(b: &[u8; 64]) -> G1Point {
    if b.iter().all(|&x| x == 0) {
        return G1Point::point_at_infinity(); // curve.rs:34-39
    }
    let x = BigInt::from_bytes_le(Sign::Plus, &b[..32]);
    let y = BigInt::from_bytes_le(Sign::Plus, &b[32..]);
    G1Point::new(Fq::from_value(x), Fq::from_value(y))
}

fn fq_to_le(x: &Fq) -> [u8; 32] {
    let (_, digits) = x.sanitize().get_value().to_u64_digits();
    let mut out = [0u8; 32];
    for (i, d) in digits.iter().take(4).enumerate() {
        out[8 * i..8 * i + 8].copy_from_slice(&d.to_le_bytes());
    }
    out
}

/// G2 wire form: x.c0 | x.c1 | y.c0 | y.c1 (Fq2 = Fq[u]/(u^2+1), efield.rs:95-98: `poly.coef` low -> high,
/// trailing zeros trimmed); infinity = 128 zero bytes.
fn g2_to_bytes(p: &G2Point) -> [u8; 128] {
    let mut out = [0u8; 128];
    if let (Some(x), Some(y)) = (&p.x, &p.y) {
        for (k, e) in [x, y].iter().enumerate() {
            for (j, c) in e.poly.coef.iter().take(2).enumerate() {
                out[64 * k + 32 * j..64 * k + 32 * j + 32].copy_from_slice(&fq_to_le(c));
            }
        }
    }
    out
}

fn g2_from_bytes(b: &[u8]) -> G2Point {
    if b.iter().all(|&x| x == 0) {
        return G2Point::point_at_infinity();
    }
    let fq = |s: &[u8]| Fq::from_value(BigInt::from_bytes_le(Sign::Plus, s));
    let fq2 = |s: &[u8]| Fq2::new(Polynomial { coef: vec![fq(&s[..32]), fq(&s[32..64])] });
    G2Point::new(fq2(&b[..64]), fq2(&b[64..128]))
}

/// [alpha^i] g2 for i < n, computed on the GPU (kzg.rs:37 for n = 2, kzg.rs:47-52 for n = max_d + 1)
fn g2_powers(ctx: *mut sys::myzkp_ctx, alpha: &FqOrder, g2: &G2Point, n: usize) -> Vec<G2Point> {
    let a = scalar_to_le(alpha);
    let base = g2_to_bytes(g2);
    let mut out = vec![0u8; 128 * n];
    check(ctx, unsafe { sys::myzkp_srs_generate_g2(ctx, a.as_ptr(), base.as_ptr(), 0, n, out.as_mut_ptr()) });
    out.chunks_exact(128).map(g2_from_bytes).collect()
}

fn setup(g1: &G1Point, g2: &G2Point, max_d: usize, n_g2: usize) -> GpuPublicKeyKZG {
    let alpha = FqOrder::random_element(&[]); // kzg.rs:28
    setup_with_alpha(g1, g2, max_d, n_g2, &alpha)
}

/// setup with the trapdoor injected (tests, reproducible benches; the reference draws it unseeded, kzg.rs:28)
pub fn setup_kzg_with_alpha(g1: &G1Point, g2: &G2Point, max_d: usize, alpha: &FqOrder) -> GpuPublicKeyKZG {
    setup_with_alpha(g1, g2, max_d, 2, alpha)
}

fn setup_with_alpha(g1: &G1Point, g2: &G2Point, max_d: usize, n_g2: usize, alpha: &FqOrder) -> GpuPublicKeyKZG {
    assert!(*g1 == BN128::generator_g1());
    let mut ctx = ptr::null_mut();
    let code = unsafe { sys::myzkp_ctx_create(&mut ctx, 0) };
    assert!(code == sys::MYZKP_OK, "no usable CUDA device (there is no CPU fallback)");
    let a = scalar_to_le(alpha);
    check(ctx, unsafe { sys::myzkp_srs_generate_g1(ctx, a.as_ptr(), 0, max_d + 1) });
    let powers_2 = g2_powers(ctx, alpha, g2, n_g2);
    GpuPublicKeyKZG { ctx, powers_2 }
}

/// setup_kzg (kzg.rs:27-40).  `g1` must be BN128::generator_g1(); alpha is drawn like the reference does.
pub fn setup_kzg(g1: &G1Point, g2: &G2Point, max_d: usize) -> GpuPublicKeyKZG {
    setup(g1, g2, max_d, 2) // powers_2 = [g2, [alpha]g2], kzg.rs:37
}

/// setup_kzg_with_full_g2 (kzg.rs:42-55): powers_2 = [alpha^i]g2 for i = 0..=max_d
pub fn setup_kzg_with_full_g2(g1: &G1Point, g2: &G2Point, max_d: usize) -> GpuPublicKeyKZG {
    setup(g1, g2, max_d, max_d + 1)
}

/// accumulate_curve_points over G2 (zksnark/utils.rs:83-93): sum_i assignment[i] * g_vec[i] on the GPU
pub fn accumulate_curve_points_g2(g_vec: &[G2Point], assignment: &[FqOrder], pk: &GpuPublicKeyKZG) -> G2Point {
    let n = g_vec.len().min(assignment.len()); // zip() stops at the shorter slice
    let scalars = marshal_scalars(&assignment[..n]);
    let mut pts = Vec::with_capacity(128 * n);
    for g in &g_vec[..n] {
        pts.extend_from_slice(&g2_to_bytes(g));
    }
    let mut out = [0u8; 128];
    check(pk.ctx, unsafe { sys::myzkp_g2_msm(pk.ctx, scalars.as_ptr(), pts.as_ptr(), n, out.as_mut_ptr()) });
    g2_from_bytes(&out)
}

fn g1_to_bytes(p: &G1Point) -> [u8; 64] {
    let mut out = [0u8; 64];
    if let (Some(x), Some(y)) = (&p.x, &p.y) {
        out[..32].copy_from_slice(&fq_to_le(x));
        out[32..].copy_from_slice(&fq_to_le(y));
    }
    out
}

/// prod_i e(g1[i], g2[i]) == 1 on the GPU: the Miller loops run side by side, one final exponentiation
/// (optimal_ate_pairing, curve/bn128.rs:147-181).
pub fn pairing_product_is_one(g1: &[G1Point], g2: &[G2Point], pk: &GpuPublicKeyKZG) -> bool {
    let n = g1.len().min(g2.len());
    let a: Vec<u8> = g1[..n].iter().flat_map(|p| g1_to_bytes(p)).collect();
    let b: Vec<u8> = g2[..n].iter().flat_map(|p| g2_to_bytes(p)).collect();
    let mut ok: std::os::raw::c_int = 0;
    check(pk.ctx, unsafe { sys::myzkp_pairing_product_is_one(pk.ctx, a.as_ptr(), b.as_ptr(), n, &mut ok) });
    ok != 0
}

/// verify_kzg (kzg.rs:90-102), same argument list: e(C, g2) == e(W, [alpha]g2 - [u]g2) * e(g1, g2)^y, evaluated as
/// the product e(C, g2) * e(-W, [alpha - u]g2) * e([-y]g1, g2) == 1.  g1 = powers_1[0] is read from the device
/// (kzg.rs:91); the three small group operations use the reference's own point arithmetic; the pairings run on
/// the GPU.
pub fn verify_kzg(u: &FqOrder, c: &CommitmentKZG, proof: &ProofKZG, pk: &GpuPublicKeyKZG) -> bool {
    let mut g1b = [0u8; 64];
    check(pk.ctx, unsafe { sys::myzkp_srs_read_g1(pk.ctx, 0, 1, g1b.as_mut_ptr()) });
    let g1 = &point_from_bytes(&g1b);
    let g2 = &pk.powers_2[0];
    let g2_alpha_minus_u = pk.powers_2[1].clone() - g2.mul_ref(u.clone().get_value());
    let minus_y = (FqOrder::zero() - proof.y.clone()).sanitize();
    let g1_minus_y = g1.mul_ref(minus_y.get_value());
    pairing_product_is_one(
        &[c.clone(), -proof.w.clone(), g1_minus_y],
        &[g2.clone(), g2_alpha_minus_u, g2.clone()],
        pk,
    )
}

/// commit_kzg (kzg.rs:57-59)
pub fn commit_kzg(f: &Polynomial<FqOrder>, pk: &GpuPublicKeyKZG) -> CommitmentKZG {
    let bytes = marshal_scalars(&f.coef);
    let mut out = [0u8; 64];
    check(pk.ctx, unsafe { sys::myzkp_kzg_commit(pk.ctx, bytes.as_ptr(), f.coef.len(), out.as_mut_ptr()) });
    point_from_bytes(&out)
}

/// open_kzg (kzg.rs:61-72)
pub fn open_kzg(f: &Polynomial<FqOrder>, u: &FqOrder, pk: &GpuPublicKeyKZG) -> ProofKZG {
    let bytes = marshal_scalars(&f.coef);
    let ub = scalar_to_le(u);
    let (mut y, mut w) = ([0u8; 32], [0u8; 64]);
    check(pk.ctx, unsafe {
        sys::myzkp_kzg_open(pk.ctx, bytes.as_ptr(), f.coef.len(), ub.as_ptr(), y.as_mut_ptr(), w.as_mut_ptr())
    });
    ProofKZG { y: FqOrder::from_value(BigInt::from_bytes_le(Sign::Plus, &y)), w: point_from_bytes(&w) }
}

pub struct BatchProofKZG {
    pub ys: Vec<FqOrder>,
    pub w: G1Point,
}
pub type ProofDegreeBound = G1Point;

/// batch_open_kzg (kzg.rs:74-88): ys[i] = f(us[i]) and W = commit((f - I) / prod (x - us[i])) on the GPU
pub fn batch_open_kzg(f: &Polynomial<FqOrder>, us: &[FqOrder], pk: &GpuPublicKeyKZG) -> BatchProofKZG {
    let bytes = marshal_scalars(&f.coef);
    let ub = marshal_scalars(us);
    let mut ys = vec![0u8; 32 * us.len()];
    let mut w = [0u8; 64];
    check(pk.ctx, unsafe {
        sys::myzkp_kzg_batch_open(pk.ctx, bytes.as_ptr(), f.coef.len(), ub.as_ptr(), us.len(), ys.as_mut_ptr(), w.as_mut_ptr())
    });
    BatchProofKZG {
        ys: ys.chunks_exact(32).map(|c| FqOrder::from_value(BigInt::from_bytes_le(Sign::Plus, c))).collect(),
        w: point_from_bytes(&w),
    }
}

/// prove_degree_bound (kzg.rs:121-134): an MSM of f against the SRS window starting at max_d - d
pub fn prove_degree_bound(f: &Polynomial<FqOrder>, pk: &GpuPublicKeyKZG, d: usize) -> ProofDegreeBound {
    let bytes = marshal_scalars(&f.coef);
    let mut out = [0u8; 64];
    check(pk.ctx, unsafe { sys::myzkp_kzg_prove_degree_bound(pk.ctx, bytes.as_ptr(), f.coef.len(), d, out.as_mut_ptr()) });
    point_from_bytes(&out)
}

/// verify_degree_bound (kzg.rs:136-144) as e(proof, g2) * e(-c, [alpha^(max_d - d)]g2) == 1; needs the full G2 powers
pub fn verify_degree_bound(c: &CommitmentKZG, proof: &ProofDegreeBound, pk: &GpuPublicKeyKZG, d: usize) -> bool {
    let max_d = unsafe { sys::myzkp_srs_len(pk.ctx) } - 1;
    pairing_product_is_one(&[proof.clone(), -c.clone()], &[pk.powers_2[0].clone(), pk.powers_2[max_d - d].clone()], pk)
}

pub struct ProofGemini {
    pub es: Vec<BatchProofKZG>,
    pub degree_proofs: Vec<ProofDegreeBound>,
}

/// split_and_fold (gemini.rs:51-103) fused with commit_gemini (gemini.rs:112-114): the folds are computed and committed
/// on the GPU; returns the log2(n) + 1 commitments and the folded polynomials (without the original).
pub fn split_and_fold_commit(
    coef: &[FqOrder],
    rhos: &[FqOrder],
    pk: &GpuPublicKeyKZG,
) -> Result<(Vec<CommitmentKZG>, Vec<Polynomial<FqOrder>>), SplitFoldError> {
    let n = coef.len();
    let m = rhos.len();
    // the reference's own checks, before anything crosses the FFI (gemini.rs:55-66); the C ABI checks them again
    // through its n_rhos argument, so a wrong count can never size a buffer
    if n.count_ones() != 1 {
        return Err(SplitFoldError::CoefsNotPowerOfTwo { found: n });
    }
    let log2_n = (usize::BITS - 1 - n.leading_zeros()) as usize;
    if m != log2_n {
        return Err(SplitFoldError::PointsLenMismatch { expected: log2_n, found: m });
    }
    let (cb, rb) = (marshal_scalars(coef), marshal_scalars(rhos));
    let mut out = vec![0u8; 64 * (m + 1)];
    let mut folds = vec![0u8; 32 * n.saturating_sub(1)];
    check(pk.ctx, unsafe {
        sys::myzkp_gemini_fold_commit(pk.ctx, cb.as_ptr(), n, rb.as_ptr(), m, out.as_mut_ptr(), folds.as_mut_ptr())
    });
    let cms = out.chunks_exact(64).map(|c| point_from_bytes(c.try_into().unwrap())).collect();
    let mut polys = Vec::with_capacity(m);
    let (mut off, mut len) = (0usize, n / 2);
    while len >= 1 {
        let coef = folds[32 * off..32 * (off + len)]
            .chunks_exact(32)
            .map(|c| FqOrder::from_value(BigInt::from_bytes_le(Sign::Plus, c)))
            .collect();
        polys.push(Polynomial { coef });
        off += len;
        len /= 2;
    }
    Ok((cms, polys))
}

/// open_gemini (gemini.rs:116-144)
pub fn open_gemini(polys: &[Polynomial<FqOrder>], beta: &FqOrder, pk: &GpuPublicKeyKZG) -> ProofGemini {
    let num_polys = polys.len();
    let us = vec![beta.clone(), (FqOrder::zero() - beta.clone()).sanitize(), beta.pow(2)];
    ProofGemini {
        es: polys.iter().take(num_polys - 1).map(|p| batch_open_kzg(p, &us, pk)).collect(),
        degree_proofs: polys
            .iter()
            .enumerate()
            .map(|(i, p)| prove_degree_bound(p, pk, 2_usize.pow((num_polys - i - 1) as u32)))
            .collect(),
    }
}

/// commit_gemini (gemini.rs:112-114): one batched call (small polynomials run concurrently on the GPU)
pub fn commit_gemini(polys: &[Polynomial<FqOrder>], pk: &GpuPublicKeyKZG) -> Vec<CommitmentKZG> {
    let bufs: Vec<Vec<u8>> = polys.iter().map(|p| marshal_scalars(&p.coef)).collect();
    let ptrs: Vec<*const u8> = bufs.iter().map(|b| b.as_ptr()).collect();
    let lens: Vec<usize> = polys.iter().map(|p| p.coef.len()).collect();
    let mut out = vec![0u8; 64 * polys.len()];
    check(pk.ctx, unsafe {
        sys::myzkp_kzg_commit_batch(pk.ctx, ptrs.as_ptr(), lens.as_ptr(), polys.len(), out.as_mut_ptr())
    });
    out.chunks_exact(64).map(|c| point_from_bytes(c.try_into().unwrap())).collect()
}

/// Range-sharded prover, one `GpuPublicKeyKZG` per GPU: rank g holds powers_1[first, first + count).
/// `alpha` is shared by the ranks (the reference draws it inside setup_kzg, kzg.rs:28).
pub fn setup_kzg_range(device: i32, alpha: &FqOrder, first: usize, count: usize, g2: &G2Point) -> GpuPublicKeyKZG {
    let mut ctx = ptr::null_mut();
    let code = unsafe { sys::myzkp_ctx_create(&mut ctx, device) };
    assert!(code == sys::MYZKP_OK, "no usable CUDA device (there is no CPU fallback)");
    let a = scalar_to_le(alpha);
    check(ctx, unsafe { sys::myzkp_srs_generate_g1(ctx, a.as_ptr(), first, count) });
    let powers_2 = g2_powers(ctx, alpha, g2, 2);
    GpuPublicKeyKZG { ctx, powers_2 }
}

/// Map every rank's exchange buffer into every other rank (all ranks in this process).
pub fn attach_peers(ranks: &[&GpuPublicKeyKZG]) {
    let ctxs: Vec<*mut sys::myzkp_ctx> = ranks.iter().map(|r| r.ctx).collect();
    for r in ranks {
        check(r.ctx, unsafe { sys::myzkp_peer_export(r.ctx, ptr::null_mut()) });
    }
    for (g, r) in ranks.iter().enumerate() {
        check(r.ctx, unsafe { sys::myzkp_peer_attach_local(r.ctx, g as i32, ranks.len() as i32, ctxs.as_ptr()) });
    }
}

/// commit_kzg of the whole polynomial from this rank's coefficient slice; every rank returns the same point.
/// Blocks until the peers have called it too: drive each rank from its own thread.
pub fn commit_kzg_sharded(local_slice: &[FqOrder], pk: &GpuPublicKeyKZG) -> CommitmentKZG {
    let bytes = marshal_scalars(local_slice);
    let mut out = [0u8; 64];
    check(pk.ctx, unsafe { sys::myzkp_kzg_commit_sharded(pk.ctx, bytes.as_ptr(), local_slice.len(), out.as_mut_ptr()) });
    point_from_bytes(&out)
}

/// open_kzg of the whole polynomial from this rank's coefficient slice (every rank returns the same proof)
pub fn open_kzg_sharded(local_slice: &[FqOrder], u: &FqOrder, pk: &GpuPublicKeyKZG) -> ProofKZG {
    let bytes = marshal_scalars(local_slice);
    let ub = scalar_to_le(u);
    let (mut y, mut w) = ([0u8; 32], [0u8; 64]);
    check(pk.ctx, unsafe {
        sys::myzkp_kzg_open_sharded(pk.ctx, bytes.as_ptr(), local_slice.len(), ub.as_ptr(), y.as_mut_ptr(), w.as_mut_ptr())
    });
    ProofKZG { y: FqOrder::from_value(BigInt::from_bytes_le(Sign::Plus, &y)), w: point_from_bytes(&w) }
}

/// One process, several GPUs (csrc/multi.cu): the SRS is range-sharded over `devices`, commit / open split the
/// coefficients, every GPU runs its range and the partials meet in the peer-memory exchange kernel.  One call from
/// one host thread; the library drives the devices from its own threads.
pub struct MultiGpuKZG {
    m: *mut sys::myzkp_mctx,
    pub powers_2: Vec<G2Point>,
}
unsafe impl Send for MultiGpuKZG {}

impl Drop for MultiGpuKZG {
    fn drop(&mut self) {
        unsafe { sys::myzkp_mctx_destroy(self.m) };
    }
}

fn mcheck(m: *mut sys::myzkp_mctx, code: i32) {
    if code != sys::MYZKP_OK {
        let msg = unsafe { CStr::from_ptr(sys::myzkp_mctx_last_error(m)) }.to_string_lossy().into_owned();
        panic!("myzkp_b200 error {}: {}", code, msg);
    }
}

impl MultiGpuKZG {
    fn create(devices: &[i32]) -> *mut sys::myzkp_mctx {
        let mut m = ptr::null_mut();
        let code = unsafe { sys::myzkp_mctx_create(&mut m, devices.as_ptr(), devices.len() as i32) };
        assert!(code == sys::MYZKP_OK, "no usable CUDA devices (there is no CPU fallback)");
        m
    }
    /// setup_kzg (kzg.rs:27-40) over several GPUs, trapdoor injected
    pub fn setup_with_alpha(devices: &[i32], g2: &G2Point, max_d: usize, alpha: &FqOrder) -> Self {
        let m = Self::create(devices);
        let a = scalar_to_le(alpha);
        mcheck(m, unsafe { sys::myzkp_mctx_srs_generate_g1(m, a.as_ptr(), max_d + 1) });
        let powers_2 = g2_powers(unsafe { sys::myzkp_mctx_rank(m, 0) }, alpha, g2, 2);
        MultiGpuKZG { m, powers_2 }
    }
    /// an existing reference key, sharded over the devices
    pub fn from_reference(devices: &[i32], pk: &PublicKeyKZG) -> Self {
        let m = Self::create(devices);
        let mut bytes = Vec::with_capacity(64 * pk.powers_1.len());
        for p in &pk.powers_1 {
            bytes.extend_from_slice(&g1_to_bytes(p));
        }
        mcheck(m, unsafe { sys::myzkp_mctx_srs_load_g1(m, bytes.as_ptr(), pk.powers_1.len()) });
        MultiGpuKZG { m, powers_2: pk.powers_2.clone() }
    }
    /// commit_kzg (kzg.rs:57-59)
    pub fn commit_kzg(&self, f: &Polynomial<FqOrder>) -> CommitmentKZG {
        let bytes = marshal_scalars(&f.coef);
        let mut out = [0u8; 64];
        mcheck(self.m, unsafe { sys::myzkp_mctx_kzg_commit(self.m, bytes.as_ptr(), f.coef.len(), out.as_mut_ptr()) });
        point_from_bytes(&out)
    }
    /// open_kzg (kzg.rs:61-72)
    pub fn open_kzg(&self, f: &Polynomial<FqOrder>, u: &FqOrder) -> ProofKZG {
        let bytes = marshal_scalars(&f.coef);
        let ub = scalar_to_le(u);
        let (mut y, mut w) = ([0u8; 32], [0u8; 64]);
        mcheck(self.m, unsafe {
            sys::myzkp_mctx_kzg_open(self.m, bytes.as_ptr(), f.coef.len(), ub.as_ptr(), y.as_mut_ptr(), w.as_mut_ptr())
        });
        ProofKZG { y: FqOrder::from_value(BigInt::from_bytes_le(Sign::Plus, &y)), w: point_from_bytes(&w) }
    }
}

/// The reference's exact signatures (kzg.rs:27,57,61,90; gemini.rs:112) over the reference's own `PublicKeyKZG`:
/// `use myzkp_b200::dropin::{setup_kzg, commit_kzg, open_kzg, verify_kzg}` instead of
/// `use myzkp::modules::algebra::kzg::{...}` and nothing else changes at the call sites.  The device side of a key
/// (context + resident table) is cached per key: the first call with a given `PublicKeyKZG` uploads powers_1
/// (`From<&PublicKeyKZG>`), later calls find it by the key's address, length and end points.
pub mod dropin {
    use super::*;
    use myzkp::modules::algebra::kzg::ProofKZG as RefProofKZG;
    use std::sync::{Arc, Mutex, OnceLock};

    #[derive(PartialEq, Clone)]
    struct Fingerprint {
        addr: usize,
        len: usize,
        first: [u8; 64],
        last: [u8; 64],
    }
    fn fingerprint(pk: &PublicKeyKZG) -> Fingerprint {
        let n = pk.powers_1.len();
        Fingerprint {
            addr: pk.powers_1.as_ptr() as usize,
            len: n,
            first: if n > 1 { g1_to_bytes(&pk.powers_1[1]) } else { [0u8; 64] },
            last: if n > 0 { g1_to_bytes(&pk.powers_1[n - 1]) } else { [0u8; 64] },
        }
    }
    struct Shared(Arc<Mutex<GpuPublicKeyKZG>>);
    fn cache() -> &'static Mutex<Vec<(Fingerprint, Arc<Mutex<GpuPublicKeyKZG>>)>> {
        static C: OnceLock<Mutex<Vec<(Fingerprint, Arc<Mutex<GpuPublicKeyKZG>>)>>> = OnceLock::new();
        C.get_or_init(|| Mutex::new(Vec::new()))
    }
    fn device_key(pk: &PublicKeyKZG) -> Shared {
        let fp = fingerprint(pk);
        let mut c = cache().lock().unwrap();
        if let Some((_, g)) = c.iter().find(|(f, _)| *f == fp) {
            return Shared(g.clone());
        }
        let g = Arc::new(Mutex::new(GpuPublicKeyKZG::from(pk)));
        if c.len() >= 4 {
            c.remove(0); // a handful of keys at most stay resident
        }
        c.push((fp, g.clone()));
        Shared(g)
    }

    /// kzg.rs:27-40: the powers are generated on the GPU and read back into the reference's struct; the device
    /// side stays resident for the calls below.
    pub fn setup_kzg(g1: &G1Point, g2: &G2Point, max_d: usize) -> PublicKeyKZG {
        let gpu = super::setup_kzg(g1, g2, max_d);
        let pk = gpu.to_reference();
        cache().lock().unwrap().push((fingerprint(&pk), Arc::new(Mutex::new(gpu))));
        pk
    }
    /// kzg.rs:57-59
    pub fn commit_kzg(f: &Polynomial<FqOrder>, pk: &PublicKeyKZG) -> CommitmentKZG {
        let k = device_key(pk);
        let g = k.0.lock().unwrap();
        super::commit_kzg(f, &g)
    }
    /// kzg.rs:61-72
    pub fn open_kzg(f: &Polynomial<FqOrder>, u: &FqOrder, pk: &PublicKeyKZG) -> RefProofKZG {
        let k = device_key(pk);
        let g = k.0.lock().unwrap();
        let p = super::open_kzg(f, u, &g);
        RefProofKZG { y: p.y, w: p.w }
    }
    /// kzg.rs:90-102
    pub fn verify_kzg(u: &FqOrder, c: &CommitmentKZG, proof: &RefProofKZG, pk: &PublicKeyKZG) -> bool {
        let k = device_key(pk);
        let g = k.0.lock().unwrap();
        super::verify_kzg(u, c, &ProofKZG { y: proof.y.clone(), w: proof.w.clone() }, &g)
    }
    /// gemini.rs:112-114
    pub fn commit_gemini(polys: &[Polynomial<FqOrder>], pk: &PublicKeyKZG) -> Vec<CommitmentKZG> {
        let k = device_key(pk);
        let g = k.0.lock().unwrap();
        super::commit_gemini(polys, &g)
    }
}
