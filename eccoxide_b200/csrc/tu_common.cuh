// tu_common.cuh — what every kernel TU needs: bodies, context, the batch-inversion launcher.
#pragma once
#include "host_ctx.h"
#include "kernels.cuh"
#include "kernels2.cuh"
using namespace ecb;

template <class FT, class FIN>
__global__ void __launch_bounds__(ECB_TPB) k_batch_inv(size_t T, size_t n, const u32* planes, u32* pf, FIN fin) {
    size_t t = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (t < T) batch_inv_body<FT, FIN>(t, T, n, planes, pf, fin);
}
template <class FT, class FIN>
static int launch_batch_inv(ecb_ctx* ctx, DevCtx& d, size_t n, const u32* planes, u32* pf, FIN fin, cudaStream_t s) {
    size_t T = inv_threads(d, n);
    k_batch_inv<FT, FIN><<<grid_for(T), ECB_TPB, 0, s>>>(T, n, planes, pf, fin);
    ctx->launches++;
    CU(cudaGetLastError());
    return ECB_OK;
}

