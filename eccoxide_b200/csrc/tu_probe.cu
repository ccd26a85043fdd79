// tu_probe.cu — integer-pipe (IMAD) peak probe: the denominator of every roofline fraction.
#include "host_ctx.h"
#include "dev_ops.h"
using namespace ecb;

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

