/* myzkp_b200 - C ABI of the B200-native KZG prover hot path (BN128).
 *
 * Drop-in boundary for MyZKP's `setup_kzg` / `commit_kzg` / `open_kzg` /
 * `commit_gemini` (reference: myzkp/src/modules/algebra/kzg.rs:27-72,
 * gemini.rs:51-114).  The reference has no FFI for this path (they are plain
 * generic Rust calls); its only host<->device convention is
 * myzkp/examples/sumcheck/src/utils.rs:51-72 (field element = 32 bytes,
 * little-endian, canonical, non-Montgomery), which this ABI keeps.
 *
 * Encodings
 *   Fr / Fq element : 32 B little-endian, canonical (< modulus), non-Montgomery.
 *   G1 point        : x || y (64 B); the point at infinity is 64 zero bytes
 *                     (reference: EllipticCurvePoint{x:None,y:None}, curve.rs:17-46).
 * Ownership: the caller owns every host buffer; the ctx owns all device memory.
 * Errors: 0 = ok, negative = error (message via myzkp_last_error); never aborts.
 * Threading: one in-flight call per ctx.
 * There is NO CPU fallback: every entry point runs CUDA kernels on the ctx's
 * device and fails with MYZKP_ERR_CUDA when no device is usable.
 */
#ifndef MYZKP_B200_H
#define MYZKP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct myzkp_ctx myzkp_ctx;

enum {
  MYZKP_OK = 0,
  MYZKP_ERR_INVALID_ARG = -1,  /* null pointer; n > srs_len (reference panics: polynomial.rs:162);
                                  gemini length not a power of two (gemini.rs:55-57) */
  MYZKP_ERR_NONCANONICAL = -2, /* an input limb vector >= modulus */
  MYZKP_ERR_CUDA = -3,
  MYZKP_ERR_OOM = -4,
  MYZKP_ERR_NO_SRS = -5
};

/* ---- context ---------------------------------------------------------- */
int myzkp_ctx_create(myzkp_ctx** out, int device_id);
int myzkp_ctx_destroy(myzkp_ctx* ctx);
/* Run all work of this ctx on an existing CUDA stream (cudaStream_t). */
int myzkp_ctx_set_stream(myzkp_ctx* ctx, void* cuda_stream);
/* Waits for the ctx stream.  Also reports what the asynchronous device-pointer calls could not: a scalar >= r
 * seen since the last synchronising call (MYZKP_ERR_NONCANONICAL - the flag is sticky and cleared once reported)
 * and a peer exchange that timed out (MYZKP_ERR_CUDA). */
int myzkp_ctx_sync(myzkp_ctx* ctx);
/* Size every scratch buffer for commits / opens / Gemini folds of up to n_max coefficients now (runs the
 * pipelines once on zeros), so later calls allocate nothing.  Required before sharded calls on contexts that share
 * ONE device with an attached peer (myzkp_peer_attach_local): there a later cudaMalloc would wait for the whole
 * device, i.e. for a peer's exchange kernel that is itself waiting for this rank; such contexts therefore refuse to
 * grow their scratch (MYZKP_ERR_CUDA with an explanatory message) instead of stalling.  Call again after
 * myzkp_ctx_set_msm_params or a new SRS. */
int myzkp_ctx_reserve(myzkp_ctx* ctx, size_t n_max);
const char* myzkp_last_error(const myzkp_ctx* ctx);
/* Number of this library's kernels launched by the ctx so far. */
uint64_t myzkp_kernel_launches(const myzkp_ctx* ctx);
/* MSM tuning knobs (0 = automatic): window bits c out of the windows the resident
 * table supports (4, 8, 12, 16, 20, 22, 24; or 8, 16, 24 for a very large SRS) -
 * other values fall back to automatic; entries per accumulate segment. */
int myzkp_ctx_set_msm_params(myzkp_ctx* ctx, int window_bits, int segment_len);

/* Bucket-accumulate variant: -1 = automatic (today: the XYZZ mixed-addition kernel), 0 = XYZZ only,
 * -2 = fused batched-affine pair sums (csrc/msm.cu msm_accumulate_baa; bit-exact, measured slower on
 * B200 - profiles/baa_r2.md), r >= 1 = r multi-pass batched-affine rounds (csrc/baa.cuh). */
int myzkp_ctx_set_baa_rounds(myzkp_ctx* ctx, int rounds);

/* Host-buffer commit/open upload a large polynomial in chunks on a copy stream while
 * earlier chunks are already being processed (each chunk is an MSM against its own SRS
 * range and the same window; the bucket sets are added and reduced once).  Chunk sizes grow
 * 4x so only the small first upload is exposed.  0 = automatic (see upload_chunks() in csrc/capi.cu),
 * 1..8 forces a chunk count. */
int myzkp_ctx_set_upload_chunks(myzkp_ctx* ctx, int chunks);

/* Per-phase CUDA-event timing of the MSM (events on the ctx stream, kept for the
 * last 32 MSMs so a timed loop can be read back after its final sync).
 * back = 0 is the most recent MSM.  out_ms = {recode, sort, accumulate, merge
 * heads, bucket reduce + tree sum} in ms (-1 when timing was off); out_info =
 * {window bits c, windows W, entries W*n, segment length, segments, buckets}. */
int myzkp_ctx_enable_phase_timing(myzkp_ctx* ctx, int on);
int myzkp_ctx_msm_phases(myzkp_ctx* ctx, int back, float out_ms[5], uint64_t out_info[6]);

/* pinned host memory for callers that want DMA-able buffers */
int myzkp_host_alloc(void** out, size_t bytes);
int myzkp_host_free(void* p);

/* ---- SRS = PublicKeyKZG.powers_1 (kzg.rs:8-11, 27-40) ------------------- */
/* Load n affine points.  Builds the resident table: row j holds 2^(b_j) P_i for the bit offsets b_j = multiples of
 * 4 and of 22 (70 rows, windows c in {4, 8, 12, 16, 20, 22, 24}); a very large SRS falls back to multiples of 8
 * (32 rows).  myzkp_srs_table_info reports what was built. */
int myzkp_srs_load_g1(myzkp_ctx* ctx, const uint8_t* affine_xy_le /* n*64 */, size_t n);
/* setup_kzg with the (unseeded, kzg.rs:28) alpha injected: points
 * [alpha^(first+i)]G for i < n; n = max_d + 1 (kzg.rs:32).  `first` lets a
 * rank generate only its shard of a range-sharded SRS. */
int myzkp_srs_generate_g1(myzkp_ctx* ctx, const uint8_t alpha_le[32], size_t first, size_t n);
/* G2 half of the public key, computed on the device and returned to the host (it is verifier-side and never
 * needed resident): out[i] = [alpha^(first+i)] base for i < n, 128 B per point = x.c0 | x.c1 | y.c0 | y.c1
 * (Fq2 = Fq[u]/(u^2+1), c0 + c1 u; 32 B little-endian canonical each; infinity = 128 zero bytes).
 * base_or_null = NULL selects BN128::generator_g2() (bn128.rs:190-205).  n = 2 gives setup_kzg's
 * powers_2 = [g2, [alpha]g2] (kzg.rs:37); n = max_d + 1 gives setup_kzg_with_full_g2's (kzg.rs:47-52). */
int myzkp_srs_generate_g2(myzkp_ctx* ctx, const uint8_t alpha_le[32], const uint8_t* base_or_null /* 128 B */, size_t first,
                          size_t n, uint8_t* out /* n*128 */);
int myzkp_srs_read_g1(myzkp_ctx* ctx, size_t off, size_t n, uint8_t* out /* n*64 */);
size_t myzkp_srs_len(const myzkp_ctx* ctx);
/* The resident table costs rows x 64 B per SRS point (70 rows by default = 70 GiB at 2^24 points on one GPU).
 * window_mask (bit c set <=> MSM window c usable, 1 <= c <= 24; 0 = automatic) restricts the NEXT
 * myzkp_srs_load_g1 / _generate_g1 to the rows those windows need - e.g. 1<<22 alone is 12 rows, (1<<16)|(1<<22)
 * is 27.  The MSM then picks among the available windows only. */
int myzkp_ctx_set_table_windows(myzkp_ctx* ctx, uint32_t window_mask);
int myzkp_srs_table_info(const myzkp_ctx* ctx, int* out_rows, uint64_t* out_bytes, uint32_t* out_windows);

/* ---- commit / open, host buffers --------------------------------------- */
/* commit_kzg (kzg.rs:57-59) = Polynomial::eval_with_powers_on_curve
 * (polynomial.rs:156-165): C = sum_i coef_i * powers_1[i].  n == 0 -> infinity. */
int myzkp_kzg_commit(myzkp_ctx* ctx, const uint8_t* coefs_le /* n*32 */, size_t n, uint8_t out_c[64]);
/* open_kzg (kzg.rs:61-72): y = f(u) (polynomial.rs:120-128),
 * q = (f - y)/(x - u) (polynomial.rs:371-405), W = commit(q).
 * n <= 1 -> W = infinity, y = f_0 (or 0) (polynomial.rs:372-374). */
int myzkp_kzg_open(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, const uint8_t u_le[32],
                   uint8_t out_y[32], uint8_t out_w[64]);
/* commit_gemini (gemini.rs:112-114) over k polynomials: out = k x 64 B in caller order.  Polynomials below
 * 2^19 coefficients share ONE MSM pipeline (a bucket range per polynomial), so many small commitments -
 * Gemini's folds, the rows of a DAS grid (das/avail.rs:96) - cost their entries rather than one
 * latency-bound pipeline each; larger ones get an MSM of their own.  Any error rejects the whole batch. */
int myzkp_kzg_commit_batch(myzkp_ctx* ctx, const uint8_t* const* coefs, const size_t* ns, size_t k,
                           uint8_t* out /* k*64 */);
/* split_and_fold (gemini.rs:51-103) + commit_gemini: n_pow2 = 2^m coefficients,
 * m challenges; writes m+1 commitments (original first).  If out_folds is not
 * NULL it receives the m folded polynomials back to back (2^(m-1) + ... + 1
 * coefficients, 32 B each). */
int myzkp_gemini_fold_commit(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n_pow2,
                             const uint8_t* rhos_le /* n_rhos*32 */, size_t n_rhos /* must be m = log2 n_pow2:
                             SplitFoldError::PointsLenMismatch otherwise (gemini.rs:60-66) */,
                             uint8_t* out /* (m+1)*64 */, uint8_t* out_folds /* (n_pow2-1)*32 or NULL */);
/* batch_open_kzg (kzg.rs:74-88): ys[i] = f(us[i]); W = commit((f - I)/Z) with I the
 * interpolant of (us, ys) and Z = prod (x - us[i]).  Since deg I < k the quotient is the
 * floor quotient of f by Z, computed as k successive (x - u_i) divisions.  k <= 64;
 * the us must be distinct (the reference's interpolate divides by their differences). */
int myzkp_kzg_batch_open(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, const uint8_t* us_le /* k*32 */, size_t k,
                         uint8_t* out_ys /* k*32 */, uint8_t out_w[64]);
/* prove_degree_bound (kzg.rs:121-134): commitment to f * x^(max_d - d), i.e. an MSM of
 * f against the SRS window starting at max_d - d (max_d = srs_len - 1).  deg f > d is
 * an error (the reference index-panics). */
int myzkp_kzg_prove_degree_bound(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, size_t d, uint8_t out_p[64]);
/* G1 MSM sum_i scalars[i] * points[i].  points == NULL: against the resident SRS (same as
 * myzkp_kzg_commit).  Otherwise n caller-supplied affine points (the
 * accumulate_curve_points / eval_with_powers_on_curve call sites outside KZG,
 * zksnark/utils.rs:83-93): classic windowed Pippenger over the points as given - per-window
 * buckets, then Horner over the windows (~254 dependent doublings, about 1 ms whatever n is).
 * No table of multiples is built and the resident SRS is neither used nor disturbed; a
 * coordinate >= p or a scalar >= r is MYZKP_ERR_NONCANONICAL. */
int myzkp_g1_msm(myzkp_ctx* ctx, const uint8_t* scalars_le, const uint8_t* points_or_null /* n*64 */, size_t n,
                 uint8_t out[64]);
/* G2 MSM sum_i scalars[i] * points[i] with caller-supplied affine G2 points (128 B each, layout as in
 * myzkp_srs_generate_g2): the accumulate_curve_points call sites over G2 (zksnark/utils.rs:83-93, e.g.
 * tutorial_snark/protocol_2.rs:68).  One double-and-add per term plus a tree sum - sized for the hundreds of
 * terms those callers have, not a Pippenger.  n == 0 -> infinity. */
int myzkp_g2_msm(myzkp_ctx* ctx, const uint8_t* scalars_le, const uint8_t* points /* n*128 */, size_t n, uint8_t out[128]);
/* Optimal ate pairings e(g1[i], g2[i]) - optimal_ate_pairing, curve/bn128.rs:147-181 - one warp each.
 * out[i] = the Fq12 value as the reference represents it: 12 coefficients of w^k in Fq[w]/(w^12 - 18 w^6 + 82),
 * 32 B little-endian canonical each (384 B per pairing); a point at infinity on either side gives 1. */
int myzkp_pairing(myzkp_ctx* ctx, const uint8_t* g1 /* n*64 */, const uint8_t* g2 /* n*128 */, size_t n, uint8_t* out /* n*384 */);
/* *out_is_one = (prod_i e(g1[i], g2[i]) == 1): n Miller loops side by side, ONE final exponentiation.
 * The pairing equations of verify_kzg (kzg.rs:90-102), batch_verify_kzg (kzg.rs:104-119) and
 * verify_degree_bound (kzg.rs:136-144) are such products after moving one side over with a negated point. */
int myzkp_pairing_product_is_one(myzkp_ctx* ctx, const uint8_t* g1 /* n*64 */, const uint8_t* g2 /* n*128 */, size_t n,
                                 int* out_is_one);
/* Polynomial::eval (polynomial.rs:120-128). */
int myzkp_fr_eval(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, const uint8_t u_le[32], uint8_t out_y[32]);
/* y and the quotient coefficients themselves (n-1 of them; n >= 1). */
int myzkp_fr_quotient(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n, const uint8_t u_le[32],
                      uint8_t out_y[32], uint8_t* out_q /* (n-1)*32 */);

/* ---- device-pointer variants (inputs/outputs already in HBM) ------------ */
/* Asynchronous on the ctx stream.  d_out_* must be device memory. */
int myzkp_kzg_commit_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n, void* d_out_c64);
int myzkp_kzg_open_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n, const uint8_t u_le[32],
                       void* d_out_y32, void* d_out_w64);
/* Partial MSM against the resident SRS points [srs_off, srs_off+n): leaves a
 * 128-byte XYZZ group element (internal Montgomery form) in d_out_xyzz128.
 * Used by the range-sharded multi-GPU commit: one partial per rank, exchanged
 * over NCCL and summed with myzkp_g1_sum_partials_dev. */
int myzkp_g1_msm_partial_dev(myzkp_ctx* ctx, const void* d_scalars, size_t n, size_t srs_off,
                             void* d_out_xyzz128);
/* Same with the scalars in host memory (pinned for full speed): uploads them through the chunked
 * pipeline, asynchronously on the ctx stream, and leaves the XYZZ partial on the device.  The host
 * buffer must stay valid until the stream has been synchronised. */
int myzkp_g1_msm_partial(myzkp_ctx* ctx, const uint8_t* scalars_le, size_t n, size_t srs_off, void* d_out_xyzz128);
/* ---- exchange fused over peer memory (NVLink / NVSwitch), one process per GPU -----------------------
 * Each rank owns a small exchange buffer that every peer maps (CUDA IPC).  The sharded entry points
 * below end in ONE kernel that stores this rank's partial into all peers' buffers, waits for theirs
 * and finishes (sum + affine, or scan-carry composition) in the same launch; no collective library
 * call is on the path.  Set-up: every rank calls myzkp_peer_export, the ranks exchange the 64-byte
 * handles out of band (e.g. torch.distributed.all_gather), every rank calls myzkp_peer_attach with
 * all of them.  All ranks must then issue the same sequence of sharded calls.  A rank that waits
 * longer than the timeout (default 10 s) reports MYZKP_ERR_CUDA from the next myzkp_ctx_sync. */
#define MYZKP_PEER_HANDLE_BYTES 64
int myzkp_peer_export(myzkp_ctx* ctx, uint8_t out_handle[MYZKP_PEER_HANDLE_BYTES]);
int myzkp_peer_attach(myzkp_ctx* ctx, int rank, int world, const uint8_t* handles /* world * 64 B */);
/* same for contexts living in ONE process (one host thread driving several GPUs or streams) */
int myzkp_peer_attach_local(myzkp_ctx* ctx, int rank, int world, myzkp_ctx* const* ctxs);
int myzkp_peer_detach(myzkp_ctx* ctx);
int myzkp_peer_set_timeout_ms(myzkp_ctx* ctx, uint32_t ms);
/* commit_kzg (kzg.rs:57-59) of a polynomial range-sharded over the attached ranks: this rank holds
 * SRS points [lo, lo + n_local) and the matching coefficient slice; every rank ends with the same
 * 64-byte commitment.  _dev: device pointers, asynchronous on the ctx stream. */
int myzkp_kzg_commit_sharded_dev(myzkp_ctx* ctx, const void* d_scalars, size_t n_local, void* d_out_c64);
int myzkp_kzg_commit_sharded(myzkp_ctx* ctx, const uint8_t* scalars_le, size_t n_local, uint8_t out_c[64]);
/* open_kzg over the same sharding with this rank's coefficient slice in host memory; every rank ends with (y, W). */
int myzkp_kzg_open_sharded(myzkp_ctx* ctx, const uint8_t* coefs_le, size_t n_local, const uint8_t u_le[32],
                           uint8_t out_y[32], uint8_t out_w[64]);
/* the exchange + sum + affine kernel on its own: this rank's XYZZ partial (device) -> the sum over ranks */
int myzkp_g1_exchange_sum_dev(myzkp_ctx* ctx, const void* d_partial_xyzz128, void* d_out_c64);
/* open_kzg (kzg.rs:61-72) over the same sharding: range evaluation, exchange + carry composition,
 * local quotient scan, local MSM, exchange + sum.  Every rank ends with (y, W). */
int myzkp_kzg_open_sharded_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n_local, const uint8_t u_le[32],
                               void* d_out_y32, void* d_out_w64);
/* Sum k XYZZ partials (k*128 B, device) -> canonical affine 64 B (device). */
int myzkp_g1_sum_partials_dev(myzkp_ctx* ctx, const void* d_partials, size_t k, void* d_out_c64);
/* Sharded open: per-rank pieces of the quotient scan over a contiguous
 * coefficient range f[lo, lo+n).  eval gives (h, u^n) with h = sum f_{lo+i} u^i.
 * quotient takes the carry c_n entering from above (the combined h of all
 * higher ranges) and, with c_i = f_{lo+i} + u c_{i+1}, writes
 * d_q[i] = c_{i+1} = q_{lo+i} (the quotient coefficient that pairs with SRS
 * point lo+i) and *d_c0 = c_0 (y when lo == 0; the carry for the range below
 * otherwise). */
int myzkp_fr_range_eval_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n, const uint8_t u_le[32],
                            void* d_out_h32, void* d_out_upow32);
int myzkp_fr_range_quotient_dev(myzkp_ctx* ctx, const void* d_coefs, size_t n, const uint8_t u_le[32],
                                const uint8_t carry_in_le[32], void* d_q /* n*32 */, void* d_c0 /* 32 */);

/* ---- one process, several GPUs (csrc/multi.cu) ---------------------------------------------------------
 * SURVEY 8(b)'s `myzkp_ctx_create(out, device_ids, n_dev)`: a multi-device context owns one myzkp_ctx and one host
 * thread per listed device; the SRS is range-sharded over them (contiguous ceil-split), commit / open hand every
 * device its coefficient slice and finish in the peer-memory exchange kernel.  Nothing but this library is on the
 * path (no torch.distributed, no NCCL) - it is what a single-process Rust host drives (kzg.rs:57-72).  A device id
 * may be listed more than once (several ranks on one GPU: used by the single-GPU tests). */
typedef struct myzkp_mctx myzkp_mctx;
int myzkp_device_count(void); /* usable CUDA devices (0 when there is none) */
int myzkp_mctx_create(myzkp_mctx** out, const int* device_ids, int n_dev);
int myzkp_mctx_destroy(myzkp_mctx* m);
const char* myzkp_mctx_last_error(const myzkp_mctx* m);
int myzkp_mctx_world(const myzkp_mctx* m);
myzkp_ctx* myzkp_mctx_rank(myzkp_mctx* m, int g); /* the per-device context (owned by m) */
size_t myzkp_mctx_srs_len(const myzkp_mctx* m);
int myzkp_mctx_srs_generate_g1(myzkp_mctx* m, const uint8_t alpha_le[32], size_t n);
int myzkp_mctx_srs_load_g1(myzkp_mctx* m, const uint8_t* affine_xy_le /* n*64 */, size_t n);
int myzkp_mctx_kzg_commit(myzkp_mctx* m, const uint8_t* coefs_le /* n*32 */, size_t n, uint8_t out_c[64]);
int myzkp_mctx_kzg_open(myzkp_mctx* m, const uint8_t* coefs_le, size_t n, const uint8_t u_le[32], uint8_t out_y[32],
                        uint8_t out_w[64]);

/* ---- test hooks: batched field / group ops for parity tests -------------
 * op: 0 add, 1 sub, 2 mul, 3 inverse(a) (Fermat), 4 neg(a), 5 inverse(a) (binary GCD);
 * field: 0 = Fq, 1 = Fr.
 * Inputs/outputs canonical 32 B LE, host pointers. */
int myzkp_test_field_op(myzkp_ctx* ctx, int field, int op, const uint8_t* a, const uint8_t* b,
                        uint8_t* out, size_t n);
/* op: 0 a[i]+b[i] (mixed add), 1 double(a[i]), 2 [k]a[i] with k = the 256-bit
 * integer in the first 32 bytes of b[i] (double-and-add in XYZZ), 3 a[i]+b[i]
 * through XYZZ+XYZZ with non-trivial denominators.  Points are 64 B affine. */
int myzkp_test_g1_op(myzkp_ctx* ctx, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n);
/* Entries a group may hold for the shared-memory group sort of the MSM's bucket sort (0 = default, 13312).  Tests
 * lower it so that small inputs reach the path skewed scalars take at full size (oversize groups).  Process-wide. */
int myzkp_test_set_sort_group_cap(int cap);

#ifdef __cplusplus
}
#endif
#endif /* MYZKP_B200_H */
