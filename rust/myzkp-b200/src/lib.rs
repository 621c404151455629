//! Safe wrapper that keeps the reference's KZG surface
//! (myzkp/src/modules/algebra/kzg.rs:27-72, gemini.rs:112-114) and moves the work to the GPU.
//!
//! Marshalling follows the reference's only host<->device precedent,
//! myzkp/examples/sumcheck/src/utils.rs:51-72: canonical value -> to_u64_digits -> 32 bytes LE.
use std::ffi::CStr;
use std::ptr;

use myzkp::modules::algebra::curve::bn128::{Fq, Fq2, FqOrder, G1Point, G2Point, BN128};
use myzkp::modules::algebra::field::Field;
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
    assert!(*g1 == BN128::generator_g1());
    let alpha = FqOrder::random_element(&[]); // kzg.rs:28
    let mut ctx = ptr::null_mut();
    let code = unsafe { sys::myzkp_ctx_create(&mut ctx, 0) };
    assert!(code == sys::MYZKP_OK, "no usable CUDA device (there is no CPU fallback)");
    let a = scalar_to_le(&alpha);
    check(ctx, unsafe { sys::myzkp_srs_generate_g1(ctx, a.as_ptr(), 0, max_d + 1) });
    let powers_2 = g2_powers(ctx, &alpha, g2, n_g2);
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

/// verify_kzg (kzg.rs:90-102): e(C, g2) == e(W, [alpha]g2 - [u]g2) * e(g1, g2)^y, evaluated as the product
/// e(C, g2) * e(-W, [alpha - u]g2) * e([-y]g1, g2) == 1.  The three small group operations use the reference's
/// own point arithmetic; the pairings run on the GPU.
pub fn verify_kzg(u: &FqOrder, c: &CommitmentKZG, proof: &ProofKZG, g1: &G1Point, pk: &GpuPublicKeyKZG) -> bool {
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
pub fn split_and_fold_commit(coef: &[FqOrder], rhos: &[FqOrder], pk: &GpuPublicKeyKZG) -> (Vec<CommitmentKZG>, Vec<Polynomial<FqOrder>>) {
    let n = coef.len();
    let m = rhos.len();
    let (cb, rb) = (marshal_scalars(coef), marshal_scalars(rhos));
    let mut out = vec![0u8; 64 * (m + 1)];
    let mut folds = vec![0u8; 32 * n.saturating_sub(1)];
    // non power-of-two n / wrong challenge count come back as an error (SplitFoldError, gemini.rs:55-66)
    check(pk.ctx, unsafe {
        sys::myzkp_gemini_fold_commit(pk.ctx, cb.as_ptr(), n, rb.as_ptr(), out.as_mut_ptr(), folds.as_mut_ptr())
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
    (cms, polys)
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
