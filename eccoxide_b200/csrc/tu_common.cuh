// tu_common.cuh — what every kernel TU needs: bodies, context, the batch-inversion launcher.
#pragma once
#include "host_ctx.h"
#include "kernels.cuh"
#include "kernels2.cuh"
using namespace ecb;

// Work distribution of the persistent kernels (one thread per scalar multiplication, the grid sized to the resident
// blocks): every WARP fetches the next 32 consecutive elements from a counter in global memory (status[1], armed by
// reset_status) until the batch is exhausted.  A fixed stride idx += T leaves a tail: 2^20 elements over the 94 720
// resident threads of the P-256 kernel are 11.07 rounds, and the twelfth ran on 52 of 740 blocks — 7.7 % of the
// kernel; with the counter the last elements go to whichever warps finish first, spread over all SMs.
// Returns the first index of the chunk (>= n: done); all 32 lanes must call it together.
static __device__ __forceinline__ size_t warp_next_chunk(unsigned long long* status) {
    unsigned long long c = 0;
    if ((threadIdx.x & 31u) == 0) c = atomicAdd(status + 1, 1ull) + 1ull;
    c = __shfl_sync(0xffffffffu, c, 0);
    return (size_t)c * 32;
}

template <class FT, class FIN>
__global__ void __launch_bounds__(ECB_TPB) k_batch_inv_thread(size_t T, size_t n, const u32* planes, u32* pf, FIN fin) {
    size_t t = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (t < T) batch_inv_body<FT, FIN>(t, T, n, planes, pf, fin);
}

template <class FT, class FIN>
__global__ void __launch_bounds__(ECB_TPB) k_batch_inv_thread_ct(size_t T, size_t n, const u32* planes, u32* pf, FIN fin) {
    size_t t = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (t < T) batch_inv_body<FT, FIN, true>(t, T, n, planes, pf, fin);
}
// constant-time batch inversion (ct.cuh): per-thread chains around the fixed Fermat chain, on the caller's stream
template <class FT, class FIN>
static int launch_batch_inv_ct(ecb_ctx* ctx, DevCtx& d, size_t n, const u32* planes, u32* pf, FIN fin, cudaStream_t s) {
    size_t T = inv_threads(d, n);
    k_batch_inv_thread_ct<FT, FIN><<<grid_for(T), ECB_TPB, 0, s>>>(T, n, planes, pf, fin);
    ctx->launches++;
    CU(cudaGetLastError());
    return ECB_OK;
}

// Block-cooperative batch inversion: every thread runs the forward pass over its own few elements,
// the ECB_TPB chain totals of the block are inverted TOGETHER (shared memory, then a product scan
// over the lanes of warp 0, ONE safegcd inversion per block), and every thread runs its backward
// pass.  Short per-thread chains give many warps to hide the memory latency of the two passes, and
// the inversion count drops from one per thread to one per block.
template <class FT>
__device__ __forceinline__ void fe_shfl(typename FT::el& r, const typename FT::el& a, int src_lane) {
#pragma unroll
    for (int k = 0; k < FT::N; k++) r.v[k] = __shfl_sync(0xffffffffu, a.v[k], src_lane);
}
template <class FT>
__device__ __forceinline__ void sm_ld(typename FT::el& r, const u32* sh, int e) {
#pragma unroll
    for (int k = 0; k < FT::N; k++) r.v[k] = sh[k * ECB_TPB + e];
}
template <class FT>
__device__ __forceinline__ void sm_st(u32* sh, int e, const typename FT::el& a) {
#pragma unroll
    for (int k = 0; k < FT::N; k++) sh[k * ECB_TPB + e] = a.v[k];
}
template <class FT, class FIN>
__global__ void __launch_bounds__(ECB_TPB) k_batch_inv(size_t T, size_t n, const u32* planes, u32* pf, FIN fin) {
    typedef typename FT::el fe;
    constexpr int N = FT::N;
    constexpr int PER_LANE = ECB_TPB / 32;
    __shared__ u32 sh_tot[N * ECB_TPB];   // chain totals, then their inverses (limb-major: conflict-free)
    __shared__ u32 sh_pre[N * ECB_TPB];   // running products inside a lane's PER_LANE entries
    __shared__ u32 jump[SG_JUMP_WORDS];   // safegcd jump table (modinv.cuh), read after the __syncthreads() below
    sg_stage_jump_table(jump);
    const int tid = threadIdx.x, lane = tid & 31;
    size_t t = (size_t)blockIdx.x * ECB_TPB + tid;
    fe accA, accB, inv;
    batch_inv_forward<FT>(t, T, n, planes, pf, accA, accB);
    FT::mul(inv, accA, accB);
    sm_st<FT>(sh_tot, tid, inv);
    __syncthreads();
    if (tid < 32) {
        fe one, run, tmp;
        FT::set_one(one);
        run = one;
#pragma unroll 1
        for (int j = 0; j < PER_LANE; j++) {       // entry j*32 + lane (any partition will do: products commute)
            sm_ld<FT>(tmp, sh_tot, j * 32 + lane);
            sm_st<FT>(sh_pre, j * 32 + lane, run);
            FT::mul(run, run, tmp);
        }
        // inclusive prefix (P) and suffix (S) products over the lanes
        fe P = run, S = run, o, m;
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            fe_shfl<FT>(o, P, lane >= d ? lane - d : lane);
            FT::mul(m, P, o);
            if (lane >= d) P = m;
            fe_shfl<FT>(o, S, lane + d < 32 ? lane + d : lane);
            FT::mul(m, S, o);
            if (lane + d < 32) S = m;
        }
        fe total, tinv, exP, exS;
        fe_shfl<FT>(total, P, 31);
        FT::invert_warp(tinv, total, jump);              // the 32 lanes compute ONE inverse together (modinv.cuh)
        fe_shfl<FT>(exP, P, lane > 0 ? lane - 1 : 0);
        fe_shfl<FT>(exS, S, lane < 31 ? lane + 1 : 31);
        if (lane == 0) exP = one;
        if (lane == 31) exS = one;
        FT::mul(run, exP, exS);
        FT::mul(run, run, tinv);                   // inverse of this lane's product
#pragma unroll 1
        for (int j = PER_LANE; j-- > 0;) {
            sm_ld<FT>(tmp, sh_tot, j * 32 + lane);
            sm_ld<FT>(o, sh_pre, j * 32 + lane);
            FT::mul(m, run, o);
            sm_st<FT>(sh_tot, j * 32 + lane, m);   // inverse of entry j*32 + lane
            FT::mul(run, run, tmp);
        }
    }
    __syncthreads();
    sm_ld<FT>(inv, sh_tot, tid);
    fe invA, invB;
    FT::mul(invA, inv, accB);
    FT::mul(invB, inv, accA);
    batch_inv_backward<FT, FIN>(t, T, n, planes, pf, fin, invA, invB);
}
template <class FT, class FIN>
static void launch_batch_inv_on(ecb_ctx* ctx, DevCtx& d, size_t n, const u32* planes, u32* pf, FIN fin, cudaStream_t s) {
    size_t T = inv_threads(d, n);
    if (d.inv_block == 2 || (d.inv_block == 1 && FT::BLOCK_INV)) {
        unsigned g = grid_for(T);
        k_batch_inv<FT, FIN><<<g, ECB_TPB, 0, s>>>((size_t)g * ECB_TPB, n, planes, pf, fin);
    } else {
        k_batch_inv_thread<FT, FIN><<<grid_for(T), ECB_TPB, 0, s>>>(T, n, planes, pf, fin);
    }
    ctx->launches++;
}
template <class FT, class FIN>
static int launch_batch_inv(ecb_ctx* ctx, DevCtx& d, size_t n, const u32* planes, u32* pf, FIN fin, cudaStream_t s) {
    Slot& sl = *d.cur;
    // When chunks of a batch are pipelined over several streams the inversion kernel runs on a
    // high-priority side stream, so that its blocks are placed ahead of the remaining blocks of the
    // next chunk's scalar-multiplication kernel and the two overlap instead of queueing.
    if (ctx->opt_inv_hi && sl.hi && s == sl.stream) {
        CU(cudaEventRecord(sl.ev_a, s));
        CU(cudaStreamWaitEvent(sl.hi, sl.ev_a, 0));
        launch_batch_inv_on<FT, FIN>(ctx, d, n, planes, pf, fin, sl.hi);
        CU(cudaGetLastError());
        CU(cudaEventRecord(sl.ev_b, sl.hi));
        CU(cudaStreamWaitEvent(s, sl.ev_b, 0));
        return ECB_OK;
    }
    launch_batch_inv_on<FT, FIN>(ctx, d, n, planes, pf, fin, s);
    CU(cudaGetLastError());
    return ECB_OK;
}
