//! Raw bindings: one declaration per symbol of `include/myzkp_b200.h`.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct myzkp_ctx {
    _private: [u8; 0],
}
/// one process, several GPUs (csrc/multi.cu)
#[repr(C)]
pub struct myzkp_mctx {
    _private: [u8; 0],
}

pub const MYZKP_OK: c_int = 0;
pub const MYZKP_ERR_INVALID_ARG: c_int = -1;
pub const MYZKP_ERR_NONCANONICAL: c_int = -2;
pub const MYZKP_ERR_CUDA: c_int = -3;
pub const MYZKP_ERR_OOM: c_int = -4;
pub const MYZKP_ERR_NO_SRS: c_int = -5;

extern "C" {
    pub fn myzkp_ctx_create(out: *mut *mut myzkp_ctx, device_id: c_int) -> c_int;
    pub fn myzkp_ctx_destroy(ctx: *mut myzkp_ctx) -> c_int;
    pub fn myzkp_ctx_set_stream(ctx: *mut myzkp_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn myzkp_ctx_sync(ctx: *mut myzkp_ctx) -> c_int;
    pub fn myzkp_ctx_reserve(ctx: *mut myzkp_ctx, n_max: usize) -> c_int;
    pub fn myzkp_last_error(ctx: *const myzkp_ctx) -> *const c_char;
    pub fn myzkp_kernel_launches(ctx: *const myzkp_ctx) -> u64;
    pub fn myzkp_ctx_set_msm_params(ctx: *mut myzkp_ctx, window_bits: c_int, segment_len: c_int) -> c_int;
    pub fn myzkp_ctx_set_baa_rounds(ctx: *mut myzkp_ctx, rounds: c_int) -> c_int;
    pub fn myzkp_ctx_set_upload_chunks(ctx: *mut myzkp_ctx, chunks: c_int) -> c_int;
    pub fn myzkp_ctx_enable_phase_timing(ctx: *mut myzkp_ctx, on: c_int) -> c_int;
    pub fn myzkp_ctx_msm_phases(ctx: *mut myzkp_ctx, back: c_int, out_ms: *mut f32, out_info: *mut u64) -> c_int;
    pub fn myzkp_host_alloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn myzkp_host_free(p: *mut c_void) -> c_int;

    pub fn myzkp_srs_load_g1(ctx: *mut myzkp_ctx, affine_xy_le: *const u8, n: usize) -> c_int;
    pub fn myzkp_srs_generate_g1(ctx: *mut myzkp_ctx, alpha_le: *const u8, first: usize, n: usize) -> c_int;
    pub fn myzkp_srs_read_g1(ctx: *mut myzkp_ctx, off: usize, n: usize, out: *mut u8) -> c_int;
    pub fn myzkp_srs_generate_g2(ctx: *mut myzkp_ctx, alpha_le: *const u8, base_or_null: *const u8, first: usize, n: usize,
                                 out: *mut u8) -> c_int;
    pub fn myzkp_srs_len(ctx: *const myzkp_ctx) -> usize;
    pub fn myzkp_ctx_set_table_windows(ctx: *mut myzkp_ctx, window_mask: u32) -> c_int;
    pub fn myzkp_srs_table_info(ctx: *const myzkp_ctx, out_rows: *mut c_int, out_bytes: *mut u64, out_windows: *mut u32) -> c_int;
    pub fn myzkp_pairing(ctx: *mut myzkp_ctx, g1: *const u8, g2: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn myzkp_pairing_product_is_one(ctx: *mut myzkp_ctx, g1: *const u8, g2: *const u8, n: usize, out_is_one: *mut c_int) -> c_int;
    pub fn myzkp_g2_msm(ctx: *mut myzkp_ctx, scalars_le: *const u8, points: *const u8, n: usize, out: *mut u8) -> c_int;

    pub fn myzkp_kzg_commit(ctx: *mut myzkp_ctx, coefs_le: *const u8, n: usize, out_c: *mut u8) -> c_int;
    pub fn myzkp_kzg_open(ctx: *mut myzkp_ctx, coefs_le: *const u8, n: usize, u_le: *const u8, out_y: *mut u8,
                          out_w: *mut u8) -> c_int;
    pub fn myzkp_kzg_commit_batch(ctx: *mut myzkp_ctx, coefs: *const *const u8, ns: *const usize, k: usize,
                                  out: *mut u8) -> c_int;
    pub fn myzkp_gemini_fold_commit(ctx: *mut myzkp_ctx, coefs_le: *const u8, n_pow2: usize, rhos_le: *const u8,
                                    n_rhos: usize, out: *mut u8, out_folds: *mut u8) -> c_int;
    pub fn myzkp_kzg_batch_open(ctx: *mut myzkp_ctx, coefs_le: *const u8, n: usize, us_le: *const u8, k: usize,
                                out_ys: *mut u8, out_w: *mut u8) -> c_int;
    pub fn myzkp_kzg_prove_degree_bound(ctx: *mut myzkp_ctx, coefs_le: *const u8, n: usize, d: usize,
                                        out_p: *mut u8) -> c_int;
    pub fn myzkp_g1_msm(ctx: *mut myzkp_ctx, scalars_le: *const u8, points_or_null: *const u8, n: usize,
                        out: *mut u8) -> c_int;
    pub fn myzkp_fr_eval(ctx: *mut myzkp_ctx, coefs_le: *const u8, n: usize, u_le: *const u8, out_y: *mut u8) -> c_int;
    pub fn myzkp_fr_quotient(ctx: *mut myzkp_ctx, coefs_le: *const u8, n: usize, u_le: *const u8, out_y: *mut u8,
                             out_q: *mut u8) -> c_int;

    pub fn myzkp_kzg_commit_dev(ctx: *mut myzkp_ctx, d_coefs: *const c_void, n: usize, d_out_c64: *mut c_void) -> c_int;
    pub fn myzkp_kzg_open_dev(ctx: *mut myzkp_ctx, d_coefs: *const c_void, n: usize, u_le: *const u8,
                              d_out_y32: *mut c_void, d_out_w64: *mut c_void) -> c_int;
    pub fn myzkp_g1_msm_partial_dev(ctx: *mut myzkp_ctx, d_scalars: *const c_void, n: usize, srs_off: usize,
                                    d_out_xyzz128: *mut c_void) -> c_int;
    pub fn myzkp_g1_msm_partial(ctx: *mut myzkp_ctx, scalars_le: *const u8, n: usize, srs_off: usize,
                                d_out_xyzz128: *mut c_void) -> c_int;
    pub fn myzkp_peer_export(ctx: *mut myzkp_ctx, out_handle: *mut u8) -> c_int;
    pub fn myzkp_peer_attach(ctx: *mut myzkp_ctx, rank: c_int, world: c_int, handles: *const u8) -> c_int;
    pub fn myzkp_peer_attach_local(ctx: *mut myzkp_ctx, rank: c_int, world: c_int, ctxs: *const *mut myzkp_ctx) -> c_int;
    pub fn myzkp_peer_detach(ctx: *mut myzkp_ctx) -> c_int;
    pub fn myzkp_peer_set_timeout_ms(ctx: *mut myzkp_ctx, ms: u32) -> c_int;
    pub fn myzkp_kzg_commit_sharded_dev(ctx: *mut myzkp_ctx, d_scalars: *const c_void, n_local: usize,
                                        d_out_c64: *mut c_void) -> c_int;
    pub fn myzkp_g1_exchange_sum_dev(ctx: *mut myzkp_ctx, d_partial_xyzz128: *const c_void, d_out_c64: *mut c_void) -> c_int;
    pub fn myzkp_kzg_commit_sharded(ctx: *mut myzkp_ctx, scalars_le: *const u8, n_local: usize, out_c: *mut u8) -> c_int;
    pub fn myzkp_kzg_open_sharded(ctx: *mut myzkp_ctx, coefs_le: *const u8, n_local: usize, u_le: *const u8,
                                  out_y: *mut u8, out_w: *mut u8) -> c_int;
    pub fn myzkp_kzg_open_sharded_dev(ctx: *mut myzkp_ctx, d_coefs: *const c_void, n_local: usize, u_le: *const u8,
                                      d_out_y32: *mut c_void, d_out_w64: *mut c_void) -> c_int;
    pub fn myzkp_g1_sum_partials_dev(ctx: *mut myzkp_ctx, d_partials: *const c_void, k: usize,
                                     d_out_c64: *mut c_void) -> c_int;
    pub fn myzkp_fr_range_eval_dev(ctx: *mut myzkp_ctx, d_coefs: *const c_void, n: usize, u_le: *const u8,
                                   d_out_h32: *mut c_void, d_out_upow32: *mut c_void) -> c_int;
    pub fn myzkp_fr_range_quotient_dev(ctx: *mut myzkp_ctx, d_coefs: *const c_void, n: usize, u_le: *const u8,
                                       carry_in_le: *const u8, d_q: *mut c_void, d_c0: *mut c_void) -> c_int;

    pub fn myzkp_test_field_op(ctx: *mut myzkp_ctx, field: c_int, op: c_int, a: *const u8, b: *const u8, out: *mut u8,
                               n: usize) -> c_int;
    pub fn myzkp_test_g1_op(ctx: *mut myzkp_ctx, op: c_int, a: *const u8, b: *const u8, out: *mut u8, n: usize) -> c_int;
    pub fn myzkp_test_set_sort_group_cap(cap: c_int) -> c_int;

    pub fn myzkp_device_count() -> c_int;
    pub fn myzkp_mctx_create(out: *mut *mut myzkp_mctx, device_ids: *const c_int, n_dev: c_int) -> c_int;
    pub fn myzkp_mctx_destroy(m: *mut myzkp_mctx) -> c_int;
    pub fn myzkp_mctx_last_error(m: *const myzkp_mctx) -> *const c_char;
    pub fn myzkp_mctx_world(m: *const myzkp_mctx) -> c_int;
    pub fn myzkp_mctx_rank(m: *mut myzkp_mctx, g: c_int) -> *mut myzkp_ctx;
    pub fn myzkp_mctx_srs_len(m: *const myzkp_mctx) -> usize;
    pub fn myzkp_mctx_srs_generate_g1(m: *mut myzkp_mctx, alpha_le: *const u8, n: usize) -> c_int;
    pub fn myzkp_mctx_srs_load_g1(m: *mut myzkp_mctx, affine_xy_le: *const u8, n: usize) -> c_int;
    pub fn myzkp_mctx_kzg_commit(m: *mut myzkp_mctx, coefs_le: *const u8, n: usize, out_c: *mut u8) -> c_int;
    pub fn myzkp_mctx_kzg_open(m: *mut myzkp_mctx, coefs_le: *const u8, n: usize, u_le: *const u8, out_y: *mut u8,
                               out_w: *mut u8) -> c_int;
}
