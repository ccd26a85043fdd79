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
    Slot& sl = *d.cur;
    // The inversion kernel is a long dependent chain per thread with few blocks.  When chunks of a
    // batch are pipelined over several streams it runs on a high-priority side stream, so that its
    // blocks are placed ahead of the remaining blocks of the next chunk's scalar-multiplication
    // kernel and the two overlap instead of queueing behind each other.
    if (ctx->opt_inv_hi && sl.hi && s == sl.stream) {
        CU(cudaEventRecord(sl.ev_a, s));
        CU(cudaStreamWaitEvent(sl.hi, sl.ev_a, 0));
        k_batch_inv<FT, FIN><<<grid_for(T), ECB_TPB, 0, sl.hi>>>(T, n, planes, pf, fin);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(sl.ev_b, sl.hi));
        CU(cudaStreamWaitEvent(s, sl.ev_b, 0));
        return ECB_OK;
    }
    k_batch_inv<FT, FIN><<<grid_for(T), ECB_TPB, 0, s>>>(T, n, planes, pf, fin);
    ctx->launches++;
    CU(cudaGetLastError());
    return ECB_OK;
}

