// dev_ops.h — device-level operations (enqueue on a stream, no synchronisation), one
// translation unit per kernel family.  Called by the C ABI in eccbatch.cu.
#pragma once
#include "host_ctx.h"

int dev_ed25519_build_table(ecb_ctx* ctx, DevCtx& d, int W);
int dev_ed25519_table(ecb_ctx* ctx, DevCtx& d);   // build the comb of the configured shape if the device does not hold it
// enc_stride_words: distance between compressed outputs in 32-bit words (0 = packed, 8)
int dev_ed25519_mul_base(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, bool compressed, cudaStream_t s,
                         size_t enc_stride_words = 0);
// ct: the scalar multiplications by secrets run the constant-time kernels (ct.cuh); false = the fast variable-time comb
int dev_ed25519_mul_base_ct(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, bool compressed, cudaStream_t s,
                            size_t enc_stride_words = 0);
int dev_ed25519_public_from_seed(ecb_ctx* ctx, DevCtx& d, const unsigned char* d_seeds, size_t n, u32* d_pub, cudaStream_t s, bool ct);
int dev_ed25519_sign(ecb_ctx* ctx, DevCtx& d, const unsigned char* d_seeds, const unsigned char* d_pub, const unsigned char* d_msgs,
                     const unsigned long long* d_off, size_t n, unsigned char* d_sig, cudaStream_t s, bool ct);
int dev_ed25519_decompress(ecb_ctx* ctx, DevCtx& d, const u32* d_enc, size_t n, u32* d_out, unsigned char* d_ok, cudaStream_t s);
int dev_ristretto255_decompress(ecb_ctx* ctx, DevCtx& d, const u32* d_enc, size_t n, u32* d_out, unsigned char* d_ok, cudaStream_t s);
int dev_ristretto255_compress(ecb_ctx* ctx, DevCtx& d, const u32* d_xy, size_t n, u32* d_enc, cudaStream_t s);
int dev_ristretto255_mul(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_enc, size_t n, u32* d_out, cudaStream_t s);
int dev_ristretto255_mul_base(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, cudaStream_t s);
int dev_ed25519_mul(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_p, size_t n, u32* d_out, cudaStream_t s);
int dev_ed25519_verify(ecb_ctx* ctx, DevCtx& d, const u32* a, const u32* r, const u32* sl, const u32* kl, size_t n,
                       unsigned char* ok, cudaStream_t s);
int dev_ed25519_verify_msgs(ecb_ctx* ctx, DevCtx& d, const unsigned char* a, const unsigned char* sig, const unsigned char* d_msgs,
                            const unsigned long long* d_off, size_t n, unsigned char* ok, cudaStream_t s);
int dev_x25519_base(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, cudaStream_t s);
int dev_x25519(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_u, size_t n, u32* d_out, cudaStream_t s);
int dev_x448(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_u, size_t n, u32* d_out, cudaStream_t s);
int dev_wei_mul_p256(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_p, const unsigned char* d_inf_in, size_t n,
                     u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_wei_mul_p384(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_p, const unsigned char* d_inf_in, size_t n,
                     u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_wei_mul_bls(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_p, const unsigned char* d_inf_in, size_t n,
                    u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_wei_mul_k256(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_p, const unsigned char* d_inf_in, size_t n,
                     u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_wei_mul_base_k256(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_wei_decompress_k256(ecb_ctx* ctx, DevCtx& d, const u32* d_x, const unsigned char* d_sign, size_t n, u32* d_out, unsigned char* d_ok, cudaStream_t s);
int dev_wei_table_k256(ecb_ctx* ctx, DevCtx& d);
// multi-scalar multiplication (msm.cuh): the device's partial sum of its slice, then the sum of the partials made affine
int dev_wei_msm_bls(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_p, size_t n, u32* d_partial, cudaStream_t s);
int dev_wei_msm_finish_bls(ecb_ctx* ctx, DevCtx& d, const u32* d_partials, int np, u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_wei_msm_k256(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_p, size_t n, u32* d_partial, cudaStream_t s);
int dev_wei_msm_finish_k256(ecb_ctx* ctx, DevCtx& d, const u32* d_partials, int np, u32* d_out, unsigned char* d_inf, cudaStream_t s);
// fixed base (comb): d_out / d_inf as dev_wei_mul_*; *_table makes sure the device comb exists
int dev_wei_mul_base_p256(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_wei_mul_base_p384(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_wei_mul_base_bls(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, unsigned char* d_inf, cudaStream_t s);
// wire formats (kernels3.cuh): PointAffine::decompress per curve, BLS12-381 G1 standard encodings
int dev_wei_decompress_p256(ecb_ctx* ctx, DevCtx& d, const u32* d_x, const unsigned char* d_sign, size_t n, u32* d_out, unsigned char* d_ok, cudaStream_t s);
int dev_wei_decompress_p384(ecb_ctx* ctx, DevCtx& d, const u32* d_x, const unsigned char* d_sign, size_t n, u32* d_out, unsigned char* d_ok, cudaStream_t s);
int dev_wei_decompress_bls(ecb_ctx* ctx, DevCtx& d, const u32* d_x, const unsigned char* d_sign, size_t n, u32* d_out, unsigned char* d_ok, cudaStream_t s);
int dev_bls_g1_from_compressed(ecb_ctx* ctx, DevCtx& d, const u32* d_enc, size_t n, int check, u32* d_out, unsigned char* d_ok, cudaStream_t s);
int dev_bls_g1_to_compressed(ecb_ctx* ctx, DevCtx& d, const u32* d_xy, const unsigned char* d_inf, size_t n, u32* d_enc, cudaStream_t s);
int dev_bls_g1_from_uncompressed(ecb_ctx* ctx, DevCtx& d, const u32* d_enc, size_t n, int check, u32* d_out, unsigned char* d_inf,
                                 unsigned char* d_ok, cudaStream_t s);
int dev_bls_g1_to_uncompressed(ecb_ctx* ctx, DevCtx& d, const u32* d_xy, const unsigned char* d_inf, size_t n, u32* d_enc, cudaStream_t s);
int dev_wei_table_p256(ecb_ctx* ctx, DevCtx& d);
int dev_wei_table_p384(ecb_ctx* ctx, DevCtx& d);
int dev_wei_table_bls(ecb_ctx* ctx, DevCtx& d);
int dev_ecdsa_p256(ecb_ctx* ctx, DevCtx& d, const u32* d_q, const u32* d_z, const u32* d_rs, size_t n, unsigned char* d_ok,
                   cudaStream_t s);
int dev_ecdsa_p384(ecb_ctx* ctx, DevCtx& d, const u32* d_q, const u32* d_z, const u32* d_rs, size_t n, unsigned char* d_ok,
                   cudaStream_t s);
int dev_ecdsa_msgs_p256(ecb_ctx* ctx, DevCtx& d, const u32* d_q, const unsigned char* d_msgs, const unsigned long long* d_off, int hash,
                        const u32* d_rs, size_t n, unsigned char* d_ok, cudaStream_t s);
int dev_ecdsa_msgs_p384(ecb_ctx* ctx, DevCtx& d, const u32* d_q, const unsigned char* d_msgs, const unsigned long long* d_off, int hash,
                        const u32* d_rs, size_t n, unsigned char* d_ok, cudaStream_t s);
// ct: constant-time k G and k^-1 (ct.cuh); false = the fast variable-time forms
int dev_wei_mul_base_ct_p256(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_wei_mul_base_ct_p384(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, unsigned char* d_inf, cudaStream_t s);
int dev_ecdsa_sign_p256(ecb_ctx* ctx, DevCtx& d, const u32* d_d, const u32* d_k, const u32* d_z, size_t n, u32* d_rs, unsigned char* d_ok,
                        cudaStream_t s, bool ct);
int dev_ecdsa_sign_p384(ecb_ctx* ctx, DevCtx& d, const u32* d_d, const u32* d_k, const u32* d_z, size_t n, u32* d_rs, unsigned char* d_ok,
                        cudaStream_t s, bool ct);
int dev_ecdsa_sign_msgs_p256(ecb_ctx* ctx, DevCtx& d, const u32* d_d, const u32* d_k, const unsigned char* d_msgs, const unsigned long long* d_off,
                             int hash, size_t n, u32* d_rs, unsigned char* d_ok, cudaStream_t s, bool ct);
int dev_ecdsa_sign_msgs_p384(ecb_ctx* ctx, DevCtx& d, const u32* d_d, const u32* d_k, const unsigned char* d_msgs, const unsigned long long* d_off,
                             int hash, size_t n, u32* d_rs, unsigned char* d_ok, cudaStream_t s, bool ct);
int dev_imad_probe(ecb_ctx* ctx, DevCtx& d, int variant, int iters, double* macs_per_s, double* ms_out);
int dev_latency_probe(ecb_ctx* ctx, DevCtx& d, int variant, int threads, int reps, double* cycles, double* mhz);
int dev_fieldmul_probe(ecb_ctx* ctx, DevCtx& d, int num, int den, int blocks_per_sm, int reps, double* muls_per_s, double* check);
