// tu_probe.cu — integer-pipe (IMAD) peak probe: the denominator of every roofline fraction.
#include "host_ctx.h"
#include "dev_ops.h"
#include "fused.cuh"
#include "fe43.cuh"
using namespace ecb;

// field-multiplication throughput on the two multiplier pipes -----------------------------------
// Every thread runs a dependent chain of `reps` GF(2^255-19) products; warps with (warp % den) < num use the
// FP64-pipe field (fe43.cuh), the others the integer field (fe25519.cuh).  num/den = 0/1: integer only,
// 1/1: FP64 only, 1/2: every other warp.  Reports products per second over the whole chip.
__global__ void __launch_bounds__(128) k_fieldmul_probe(int reps, int num, int den, u32* sink) {
    const u32 tid = blockIdx.x * 128 + threadIdx.x;
    const int warp = threadIdx.x >> 5;
    u32 w[8], yw[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = 0x9e3779b9u * (tid + 1 + i); yw[i] = 0x85ebca6bu * (tid + 3 + i) ^ 0x1234567u; }
    u32 acc = 0;
    if ((warp % den) < num) {
        fe43 x, y;
        F43::from_words(x, w);
        F43::from_words(y, yw);
#pragma unroll 1
        for (int r = 0; r < reps; r++) F43::mul(x, x, y);
        fe25519 o;
        F43::to_fe25519(o, x);
#pragma unroll
        for (int i = 0; i < 8; i++) acc ^= o.v[i];
    } else {
        fe25519 x, y;
        F25519::from_words(x, w);
        F25519::from_words(y, yw);
#pragma unroll 1
        for (int r = 0; r < reps; r++) F25519::mul(x, x, y);
#pragma unroll
        for (int i = 0; i < 8; i++) acc ^= x.v[i];
    }
    sink[tid] = acc;
}
int dev_fieldmul_probe(ecb_ctx* ctx, DevCtx& d, int num, int den, int blocks_per_sm, int reps, double* muls_per_s, double* check) {
    if (den < 1 || num < 0 || num > den || blocks_per_sm < 1 || blocks_per_sm > 16 || reps < 1) return ECB_ERR_INVALID_ARG;
    unsigned blocks = (unsigned)(d.sm_count * blocks_per_sm);
    TRY(ensure(ctx, d.cur->aux, (size_t)blocks * 128 * sizeof(u32)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {
        CU(cudaEventRecord(e0, d.stream));
        k_fieldmul_probe<<<blocks, 128, 0, d.stream>>>(reps, num, den, (u32*)d.cur->aux.p);
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(e1, d.stream));
        CU(cudaEventSynchronize(e1));
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (muls_per_s) *muls_per_s = (double)blocks * 128.0 * reps / (ms * 1e-3);
    if (check) {   // both fields compute the same chain from the same words: XOR of thread 0's result words
        u32 v = 0;
        CU(cudaMemcpy(&v, d.cur->aux.p, sizeof(u32), cudaMemcpyDeviceToHost));
        *check = (double)v;
    }
    return ECB_OK;
}

// latency probe -------------------------------------------------------------------------------
// One block per SM; lane 0 of warp 0 reports clock64() cycles per operation of a dependent chain
// (what a latency-bound small batch pays), plus the SM clock it ran at (clock64 vs %globaltimer).
//   0: F25519::mul      1: F25519::sqr      2: F25519::invert (safegcd)      3: F25519::invert_fermat
//   4: block_invert<F25519> with the launch's block size + the inversion warp (threads <= 480)
//   5: fe_shfl_up round trip (8 shuffles)                6: ge_madd_rt     7: ge_add_p3     8: mul2 (two interleaved products)
//   9: F25519::invert_warp (one inversion by the 32 lanes of a warp; every warp of the block runs its own)
template <int V>
__global__ void __launch_bounds__(512, 1) k_latency_probe(int reps, double* out_cycles, double* out_mhz, u32* sink) {
    __shared__ u32 sh[(FUSED_MAXW + 1) * 8];
    __shared__ u32 jump[SG_JUMP_WORDS];
    sg_stage_jump_table(jump);
    __syncthreads();
    fe25519 x, y;
#pragma unroll
    for (int i = 0; i < 8; i++) { x.v[i] = 0x9e3779b9u * (threadIdx.x + 1 + i) + blockIdx.x; y.v[i] = 0x85ebca6bu * (threadIdx.x + 3 + i) ^ 0x1234567u; }
    x.v[7] &= 0x7fffffffu;
    ge_p3 P;
    ge_niels e;
    P.X = x; P.Y = y; F::set_one(P.Z); F::mul(P.T, x, y);
    e.yp = y; e.ym = x; e.t2d = P.T;
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
        if (V == 0) F::mul(x, x, y);
        else if (V == 1) F::sqr(x, x);
        else if (V == 2) { F::invert(x, x); x.v[0] ^= 5u; }
        else if (V == 3) { F::invert_fermat(x, x); x.v[0] ^= 5u; }
        else if (V == 4) { u32 z; fe25519 o; block_invert<F25519>(o, x, z, sh, jump); x = o; x.v[0] ^= 5u; }
        else if (V == 5) { fe25519 o; fe_shfl_up<F25519>(o, x, 1); x = o; x.v[0] += 1u; }
        else if (V == 6) ge_madd_rt(P, P, e, true);
        else if (V == 8) F::mul2(x, x, y, P.X, P.X, y);
        else if (V == 9) { fe25519 o; fe_shfl_idx<F25519>(o, x, 0); F::invert_warp(x, o, jump); x.v[0] ^= 5u; }
        else ge_add_p3<true>(P, P, P);
    }
    long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= x.v[i] ^ P.X.v[i] ^ P.T.v[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) {
        out_cycles[blockIdx.x] = (double)(t1 - t0) / reps;
        out_mhz[blockIdx.x] = g1 > g0 ? (double)(t1 - t0) / (double)(g1 - g0) * 1e3 : 0.0;
    }
}
int dev_latency_probe(ecb_ctx* ctx, DevCtx& d, int variant, int threads, int reps, double* cycles, double* mhz) {
    if (variant < 0 || variant > 9 || reps < 1 || threads < 32 || threads > 512 || threads % 32) return ECB_ERR_INVALID_ARG;
    unsigned blocks = (unsigned)d.sm_count;
    TRY(ensure(ctx, d.cur->aux, (size_t)blocks * 512 * sizeof(u32) + 2 * blocks * sizeof(double)));
    double* dc = (double*)d.cur->aux.p;
    double* dm = dc + blocks;
    u32* sink = (u32*)(dm + blocks);
    for (int rep = 0; rep < 2; rep++) {
        switch (variant) {
            case 0: k_latency_probe<0><<<blocks, threads, 0, d.stream>>>(reps, dc, dm, sink); break;
            case 1: k_latency_probe<1><<<blocks, threads, 0, d.stream>>>(reps, dc, dm, sink); break;
            case 2: k_latency_probe<2><<<blocks, threads, 0, d.stream>>>(reps, dc, dm, sink); break;
            case 3: k_latency_probe<3><<<blocks, threads, 0, d.stream>>>(reps, dc, dm, sink); break;
            case 4: k_latency_probe<4><<<blocks, (threads <= 480 ? threads : 480) + 32, 0, d.stream>>>(reps, dc, dm, sink); break;
            case 5: k_latency_probe<5><<<blocks, threads, 0, d.stream>>>(reps, dc, dm, sink); break;
            case 6: k_latency_probe<6><<<blocks, threads, 0, d.stream>>>(reps, dc, dm, sink); break;
            case 7: k_latency_probe<7><<<blocks, threads, 0, d.stream>>>(reps, dc, dm, sink); break;
            case 8: k_latency_probe<8><<<blocks, threads, 0, d.stream>>>(reps, dc, dm, sink); break;
            default: k_latency_probe<9><<<blocks, threads, 0, d.stream>>>(reps, dc, dm, sink); break;
        }
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(d.stream));
    }
    std::vector<double> hc(blocks), hm(blocks);
    CU(cudaMemcpy(hc.data(), dc, blocks * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hm.data(), dm, blocks * sizeof(double), cudaMemcpyDeviceToHost));
    double sc = 0, sm = 0;
    for (unsigned i = 0; i < blocks; i++) { sc += hc[i]; sm += hm[i]; }
    if (cycles) *cycles = sc / blocks;
    if (mhz) *mhz = sm / blocks;
    return ECB_OK;
}

// integer-pipe probe -----------------------------------------------------------------------
// Every variant keeps 8 independent dependency chains per thread whose multiplicand is the
// running value itself, so ptxas cannot hoist or strength-reduce the multiply (the first
// version of this probe multiplied loop-invariant operands and ptxas turned the loop into
// IADD3s).  64 counted operations per inner iteration per thread.
//   0: IMAD        a = a*x + y                     (32-bit low product)
//   1: IMAD.WIDE   acc64[i] += lo32(acc64[i+1])*y    (32x32+64 -> 64, no carry flag)
//   2: IMAD.WIDE.X two interleaved carry chains of 4 + capture = one row of mul_full<8>
//   3: IMAD.HI     a = hi(a*x) + y
//   4: DFMA        a = a*x + y   (fp64 pipe, for reference only)
//   5: IADD3       a = a + x + y (alu pipe, for reference only)
template <int V>
__global__ void __launch_bounds__(256) k_imad_probe(u32* out, int iters) {
    u32 tid = blockIdx.x * 256 + threadIdx.x;
    u32 x[8], y = (tid * 2654435761u + 12345u) | 1u;
    for (int i = 0; i < 8; i++) x[i] = ((tid + i) * 2246822519u + 7u) | 0x80000001u;
    if (V == 0 || V == 3 || V == 5) {
        u32 a[8];
        for (int i = 0; i < 8; i++) a[i] = tid + i + 1;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (V == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x[i]), "r"(y));
                    else if (V == 3) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x[i]), "r"(y));
                    else asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; xor.b32 %0, t, %2; }" : "+r"(a[i]) : "r"(x[i]), "r"(y));
                }
            }
        }
        u32 s = 0;
        for (int i = 0; i < 8; i++) s ^= a[i];
        out[tid] = s;
    } else if (V == 1) {
        unsigned long long a[8];
        for (int i = 0; i < 8; i++) a[i] = ((unsigned long long)x[i] << 32) | (tid + i + 1);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %1; mad.wide.u32 %0, lo, %2, %0; }" : "+l"(a[i]) : "l"(a[(i + 1) & 7]), "r"(y));
            }
        }
        unsigned long long s = 0;
        for (int i = 0; i < 8; i++) s ^= a[i];
        out[tid] = (u32)s ^ (u32)(s >> 32);
    } else if (V == 4) {
        double a[8], dx = 1.0 + (double)(tid & 1023) * 1e-9, dy = 1e-3;
        for (int i = 0; i < 8; i++) a[i] = 1.0 + i;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(dx), "d"(dy));
            }
        }
        double s = 0;
        for (int i = 0; i < 8; i++) s += a[i];
        out[tid] = (u32)__double2ll_rn(s);
    } else {
        // two interleaved carry chains of 4 IMAD.WIDE.U32.X + capture: exactly one row of mul_full<8>
        u32 E[10], O[10];
        for (int i = 0; i < 10; i++) { E[i] = tid + i; O[i] = tid * 3 + i; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                mac_chain<4, true>(E, x, y);
                mac_chain<4, true>(O, x + 1, y);
            }
        }
        u32 s = 0;
        for (int i = 0; i < 10; i++) s ^= E[i] ^ O[i];
        out[tid] = s;
    }
}


int dev_imad_probe(ecb_ctx* ctx, DevCtx& dref, int variant, int iters, double* macs_per_s, double* ms_out) {
    DevCtx* d = &dref;
    if (variant < 0 || variant > 5 || iters < 1) return ECB_ERR_INVALID_ARG;
    unsigned blocks = (unsigned)d->sm_count * 8;
    TRY(ensure(ctx, d->cur->aux, (size_t)blocks * 256 * sizeof(u32)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {  // first pass is warm-up
        CU(cudaEventRecord(e0, d->stream));
        switch (variant) {
            case 0: k_imad_probe<0><<<blocks, 256, 0, d->stream>>>((u32*)d->cur->aux.p, iters); break;
            case 1: k_imad_probe<1><<<blocks, 256, 0, d->stream>>>((u32*)d->cur->aux.p, iters); break;
            case 2: k_imad_probe<2><<<blocks, 256, 0, d->stream>>>((u32*)d->cur->aux.p, iters); break;
            case 3: k_imad_probe<3><<<blocks, 256, 0, d->stream>>>((u32*)d->cur->aux.p, iters); break;
            case 4: k_imad_probe<4><<<blocks, 256, 0, d->stream>>>((u32*)d->cur->aux.p, iters); break;
            case 5: k_imad_probe<5><<<blocks, 256, 0, d->stream>>>((u32*)d->cur->aux.p, iters); break;
        }
        ctx->launches++;
        CU(cudaGetLastError());
        CU(cudaEventRecord(e1, d->stream));
        CU(cudaEventSynchronize(e1));
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double macs = (double)blocks * 256.0 * (double)iters * 64.0;
    if (macs_per_s) *macs_per_s = macs / (ms * 1e-3);
    if (ms_out) *ms_out = ms;
    return ECB_OK;
}

