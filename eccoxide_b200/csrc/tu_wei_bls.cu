// tu_wei_bls.cu
#define ECB_TU_CURVE CurveBLSG1
#define ECB_TU_FN dev_wei_mul_bls
#define ECB_TU_CURVE_INDEX 2
#define ECB_TU_TABLE_FN dev_wei_table_bls
#define ECB_TU_BASE_FN dev_wei_mul_base_bls
#define ECB_TU_DECOMP_FN dev_wei_decompress_bls
#define ECB_TU_MSM_FN dev_wei_msm_bls
#define ECB_TU_MSM_FINISH_FN dev_wei_msm_finish_bls
#include "tu_wei.inc"

// ---- BLS12-381 G1 standard encodings (bls12_381/serialize.rs) ---------------------------------------
static __global__ void __launch_bounds__(ECB_TPB) k_bls_g1_from_compressed(size_t n, const u32* enc, int check, u32* out_xy,
                                                                     unsigned char* ok) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) bls_g1_from_compressed_body(idx, enc, check, out_xy, ok);
}
static __global__ void __launch_bounds__(ECB_TPB) k_bls_g1_to_compressed(size_t n, const u32* xy, const unsigned char* inf, u32* enc) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) bls_g1_to_compressed_body(idx, xy, inf, enc);
}
int dev_bls_g1_from_compressed(ecb_ctx* ctx, DevCtx& d, const u32* d_enc, size_t n, int check, u32* d_out, unsigned char* d_ok,
                               cudaStream_t s) {
    prof_mark(ctx, d, s, 0);
    k_bls_g1_from_compressed<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_enc, check, d_out, d_ok);
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);
    prof_mark(ctx, d, s, 2);
    return ECB_OK;
}
int dev_bls_g1_to_compressed(ecb_ctx* ctx, DevCtx& d, const u32* d_xy, const unsigned char* d_inf, size_t n, u32* d_enc, cudaStream_t s) {
    prof_mark(ctx, d, s, 0);
    k_bls_g1_to_compressed<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_xy, d_inf, d_enc);
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);
    prof_mark(ctx, d, s, 2);
    return ECB_OK;
}

static __global__ void __launch_bounds__(ECB_TPB) k_bls_g1_from_uncompressed(size_t n, const u32* enc, int check, u32* out_xy,
                                                                       unsigned char* inf, unsigned char* ok) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) bls_g1_from_uncompressed_body(idx, enc, check, out_xy, inf, ok);
}
static __global__ void __launch_bounds__(ECB_TPB) k_bls_g1_to_uncompressed(size_t n, const u32* xy, const unsigned char* inf, u32* enc) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) bls_g1_to_uncompressed_body(idx, xy, inf, enc);
}
int dev_bls_g1_from_uncompressed(ecb_ctx* ctx, DevCtx& d, const u32* d_enc, size_t n, int check, u32* d_out, unsigned char* d_inf,
                                 unsigned char* d_ok, cudaStream_t s) {
    k_bls_g1_from_uncompressed<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_enc, check, d_out, d_inf, d_ok);
    ctx->launches++;
    CU(cudaGetLastError());
    return ECB_OK;
}
int dev_bls_g1_to_uncompressed(ecb_ctx* ctx, DevCtx& d, const u32* d_xy, const unsigned char* d_inf, size_t n, u32* d_enc, cudaStream_t s) {
    k_bls_g1_to_uncompressed<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_xy, d_inf, d_enc);
    ctx->launches++;
    CU(cudaGetLastError());
    return ECB_OK;
}
