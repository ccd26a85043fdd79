//! Raw bindings of libeccbatch (include/eccbatch.h).  One `extern "C"` item per exported symbol.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_long, c_void};

#[repr(C)]
pub struct ecb_ctx {
    _private: [u8; 0],
}

pub const ECB_OK: c_int = 0;
pub const ECB_ERR_CUDA: c_int = -1;
pub const ECB_ERR_INVALID_ARG: c_int = -2;
pub const ECB_ERR_NONCANONICAL_SCALAR: c_int = -3;
pub const ECB_ERR_POINT_NOT_ON_CURVE: c_int = -4;
pub const ECB_ERR_OOM: c_int = -5;

pub const ECB_CURVE_P256R1: c_int = 0;
pub const ECB_CURVE_P384R1: c_int = 1;
pub const ECB_CURVE_BLS12_381_G1: c_int = 2;

#[link(name = "eccbatch")]
extern "C" {
    pub fn ecb_init(device_ids: *const c_int, n_dev: c_int, out: *mut *mut ecb_ctx) -> c_int;
    pub fn ecb_destroy(ctx: *mut ecb_ctx);
    pub fn ecb_last_error(ctx: *mut ecb_ctx) -> *const c_char;
    pub fn ecb_device_count(ctx: *mut ecb_ctx) -> c_int;
    pub fn ecb_set_option(ctx: *mut ecb_ctx, key: *const c_char, value: c_long) -> c_int;
    pub fn ecb_alloc_pinned(bytes: usize) -> *mut c_void;
    pub fn ecb_free_pinned(p: *mut c_void);

    pub fn ecb_ed25519_mul_base(ctx: *mut ecb_ctx, k_le: *const u8, n: usize, xy_le: *mut u8, bad_index: *mut usize) -> c_int;
    pub fn ecb_ed25519_mul_base_compressed(ctx: *mut ecb_ctx, k_le: *const u8, n: usize, enc: *mut u8, bad_index: *mut usize) -> c_int;
    pub fn ecb_ed25519_mul(ctx: *mut ecb_ctx, k_le: *const u8, xy_in: *const u8, n: usize, xy_out: *mut u8, bad_index: *mut usize) -> c_int;
    pub fn ecb_ed25519_verify_prehashed(ctx: *mut ecb_ctx, a_enc: *const u8, r_enc: *const u8, s_le: *const u8, k_le: *const u8, n: usize, ok: *mut u8) -> c_int;
    pub fn ecb_ed25519_verify(ctx: *mut ecb_ctx, a_enc: *const u8, msgs: *const u8, msg_off: *const u64, sig: *const u8, n: usize, ok: *mut u8) -> c_int;
    pub fn ecb_ecdsa_verify(ctx: *mut ecb_ctx, curve_id: c_int, hash: c_int, q_xy_be: *const u8, msgs: *const u8, msg_off: *const u64,
                            rs_be: *const u8, n: usize, ok: *mut u8, bad_index: *mut usize) -> c_int;
    pub fn ecb_x25519(ctx: *mut ecb_ctx, k: *const u8, u: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn ecb_x25519_base(ctx: *mut ecb_ctx, k: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn ecb_x448(ctx: *mut ecb_ctx, k: *const u8, u: *const u8, n: usize, out: *mut u8) -> c_int;
    pub fn ecb_wei_mul(ctx: *mut ecb_ctx, curve_id: c_int, k_be: *const u8, xy_be: *const u8, inf_in: *const u8, n: usize,
                       out_xy_be: *mut u8, out_inf: *mut u8, bad_index: *mut usize) -> c_int;
    pub fn ecb_wei_mul_base(ctx: *mut ecb_ctx, curve_id: c_int, k_be: *const u8, n: usize, out_xy_be: *mut u8, out_inf: *mut u8,
                            bad_index: *mut usize) -> c_int;
    pub fn ecb_ecdsa_verify_hashed(ctx: *mut ecb_ctx, curve_id: c_int, q_xy_be: *const u8, z_be: *const u8, rs_be: *const u8,
                                   n: usize, ok: *mut u8, bad_index: *mut usize) -> c_int;
    pub fn ecb_ed25519_public_from_seed(ctx: *mut ecb_ctx, seeds: *const u8, n: usize, public: *mut u8) -> c_int;
    pub fn ecb_ed25519_sign(ctx: *mut ecb_ctx, seeds: *const u8, public: *const u8, msgs: *const u8, msg_off: *const u64, n: usize,
                            sig: *mut u8) -> c_int;
    pub fn ecb_ecdsa_sign_hashed(ctx: *mut ecb_ctx, curve_id: c_int, d_be: *const u8, k_be: *const u8, z_be: *const u8, n: usize,
                                 rs_be: *mut u8, ok: *mut u8) -> c_int;
    pub fn ecb_ecdsa_sign(ctx: *mut ecb_ctx, curve_id: c_int, hash: c_int, d_be: *const u8, k_be: *const u8, msgs: *const u8,
                          msg_off: *const u64, n: usize, rs_be: *mut u8, ok: *mut u8) -> c_int;
    pub fn ecb_wei_decompress(ctx: *mut ecb_ctx, curve_id: c_int, x_be: *const u8, sign: *const u8, n: usize, out_xy_be: *mut u8,
                              ok: *mut u8) -> c_int;
    pub fn ecb_bls12_381_g1_from_compressed(ctx: *mut ecb_ctx, enc: *const u8, n: usize, check_subgroup: c_int, out_xy_be: *mut u8,
                                            ok: *mut u8) -> c_int;
    pub fn ecb_bls12_381_g1_to_compressed(ctx: *mut ecb_ctx, xy_be: *const u8, inf: *const u8, n: usize, enc: *mut u8) -> c_int;
}
