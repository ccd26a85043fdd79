// limb.cuh — 32-bit limb primitives for the sm_100a field kernels.
//
// Every multi-precision routine in this tree is written in terms of the handful
// of carry-flag primitives below.  On the device they are single PTX
// instructions; ptxas fuses each `mad.lo.cc` / `madc.hi.cc` pair into ONE SASS
// `IMAD.WIDE.U32[.X] Rd, Pcarry, Ra, Rb, Rc[, Pcarry]` (32x32+64->64 with the
// carry in a predicate), which is the instruction the whole engine is built on.
//
// ECB_HOSTSIM: tests/hostsim/ compiles the very same headers with g++ and an
// emulated carry flag so that the *sequence of primitives* (carry-chain logic,
// reductions, formulas, recoding) can be unit-tested in a container without a
// GPU.  That build is test infrastructure only: libeccbatch.so is compiled
// without ECB_HOSTSIM, every function here is then `__device__`-only and the
// library contains no host arithmetic path at all.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef ECB_HOSTSIM
#define ECB_DEV inline
#define ECB_DEVNI inline
#define ECB_UNROLL
#define ECB_NOUNROLL
#define ECB_CONST static const
#define ECB_GTABLE static const
#else
#define ECB_DEV __device__ __forceinline__
#define ECB_DEVNI __device__ __noinline__
#define ECB_UNROLL _Pragma("unroll")
#define ECB_NOUNROLL _Pragma("unroll 1")
#define ECB_CONST static __device__ __constant__ const
#define ECB_GTABLE static __device__ const   /* a table in global memory: copied to shared memory by coalesced loads */
#endif

namespace ecb {

typedef uint32_t u32;
typedef uint64_t u64;

// per-batch status word written by kernels with atomicMin: (index << 8) | code ; ~0 = no error
enum { ST_NONCANONICAL_SCALAR = 1, ST_BAD_POINT = 2 };

#ifdef ECB_HOSTSIM
// ---- emulated carry flag (tests only) ----
static thread_local u32 g_cf = 0;
ECB_DEV u32 add_cc(u32 a, u32 b) { u64 t = (u64)a + b; g_cf = (u32)(t >> 32); return (u32)t; }
ECB_DEV u32 addc_cc(u32 a, u32 b) { u64 t = (u64)a + b + g_cf; g_cf = (u32)(t >> 32); return (u32)t; }
ECB_DEV u32 addc(u32 a, u32 b) { return a + b + g_cf; }
ECB_DEV u32 sub_cc(u32 a, u32 b) { u64 t = (u64)a - b; g_cf = (u32)(t >> 63); return (u32)t; }
ECB_DEV u32 subc_cc(u32 a, u32 b) { u64 t = (u64)a - b - g_cf; g_cf = (u32)(t >> 63); return (u32)t; }
ECB_DEV u32 subc(u32 a, u32 b) { return a - b - g_cf; }
ECB_DEV u32 mul_lo(u32 a, u32 b) { return a * b; }
ECB_DEV u32 mul_hi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
ECB_DEV u32 mad_lo_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)(a * b) + c; g_cf = (u32)(t >> 32); return (u32)t; }
ECB_DEV u32 madc_lo_cc(u32 a, u32 b, u32 c) { u64 t = (u64)(u32)(a * b) + c + g_cf; g_cf = (u32)(t >> 32); return (u32)t; }
ECB_DEV u32 mad_hi_cc(u32 a, u32 b, u32 c) { u64 t = (((u64)a * b) >> 32) + c; g_cf = (u32)(t >> 32); return (u32)t; }
ECB_DEV u32 madc_hi_cc(u32 a, u32 b, u32 c) { u64 t = (((u64)a * b) >> 32) + c + g_cf; g_cf = (u32)(t >> 32); return (u32)t; }
ECB_DEV u32 madc_hi(u32 a, u32 b, u32 c) { return (u32)(((u64)a * b) >> 32) + c + g_cf; }
#else
// ---- PTX: the CC.CF flag lives between consecutive asm statements.  nvcc never
// emits .cc arithmetic of its own (64-bit adds are add.s64), and `volatile`
// keeps the statements in program order, which is how CGBN-style code works.
ECB_DEV u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ECB_DEV u32 mul_lo(u32 a, u32 b) { return a * b; }
ECB_DEV u32 mul_hi(u32 a, u32 b) { return __umulhi(a, b); }
ECB_DEV u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ECB_DEV u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ECB_DEV u32 mad_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.hi.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ECB_DEV u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ECB_DEV u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#endif

// word i of a multi-word value shifted left by s (0 < s < 32): (hi << s) | (lo >> (32 - s)), one SHF.L.W
ECB_DEV u32 shl_word(u32 lo, u32 hi, int s) {
#ifdef ECB_HOSTSIM
    return (hi << s) | (lo >> (32 - s));
#else
    return __funnelshift_l(lo, hi, s);
#endif
}

// ---------------------------------------------------------------------------
// Wide multiply-accumulate chains ("even/odd" column layout).
//
// A product a_j*b_i is a 64-bit quantity that lands on limbs (i+j, i+j+1).  All
// products with i+j even live in an "even-aligned" accumulator E, the others in
// an "odd-aligned" accumulator O; both are indexed by ABSOLUTE limb number.  One
// row (fixed b_i) is then two carry chains of IMAD.WIDE.U32.X, each touching
// consecutive aligned 64-bit slots, and the final value is E + O.
// ---------------------------------------------------------------------------

// acc[0..2K) += (a[0], a[2], ..., a[2K-2]) * bi ; the carry out of the chain is
// added into acc[2K] when CAPTURE (that word only ever holds earlier captures,
// so it cannot overflow).
template <int K, bool CAPTURE>
ECB_DEV void mac_chain(u32* acc, const u32* a, u32 bi) {
    acc[0] = mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = madc_hi_cc(a[0], bi, acc[1]);
    ECB_UNROLL
    for (int k = 1; k < K; k++) {
        acc[2 * k] = madc_lo_cc(a[2 * k], bi, acc[2 * k]);
        acc[2 * k + 1] = madc_hi_cc(a[2 * k], bi, acc[2 * k + 1]);
    }
    if (CAPTURE) acc[2 * K] = addc(acc[2 * K], 0);
}

// r[0..n) = a[0..n) + b[0..n), returns carry-out (0/1)
template <int N>
ECB_DEV u32 add_n(u32* r, const u32* a, const u32* b) {
    r[0] = add_cc(a[0], b[0]);
    ECB_UNROLL
    for (int i = 1; i < N; i++) r[i] = addc_cc(a[i], b[i]);
    return addc(0, 0);
}
// r = a - b, returns borrow-out (0/1)
template <int N>
ECB_DEV u32 sub_n(u32* r, const u32* a, const u32* b) {
    r[0] = sub_cc(a[0], b[0]);
    ECB_UNROLL
    for (int i = 1; i < N; i++) r[i] = subc_cc(a[i], b[i]);
    return subc(0, 0) & 1;
}

// Full product t[0..2N) = a[0..N) * b[0..N), N even.  N*N IMAD.WIDE + 2N-1 adds
// (+ the capture adds).
template <int N>
ECB_DEV void mul_full(u32* t, const u32* a, const u32* b) {
    u32 E[2 * N + 2], O[2 * N + 2];
    ECB_UNROLL
    for (int i = 0; i < 2 * N + 2; i++) { E[i] = 0; O[i] = 0; }
    ECB_UNROLL
    for (int i = 0; i < N; i++) {
        if ((i & 1) == 0) {
            // even row: even j -> E at limb i, odd j -> O at limb i+1
            mac_chain<N / 2, true>(E + i, a, b[i]);
            mac_chain<N / 2, true>(O + i + 1, a + 1, b[i]);
        } else {
            // odd row: odd j -> E at limb i+1, even j -> O at limb i
            mac_chain<N / 2, true>(E + i + 1, a + 1, b[i]);
            mac_chain<N / 2, true>(O + i, a, b[i]);
        }
    }
    t[0] = E[0];
    t[1] = add_cc(E[1], O[1]);
    ECB_UNROLL
    for (int i = 2; i < 2 * N; i++) t[i] = addc_cc(E[i], O[i]);
    (void)addc(0, 0);
}

// Two independent full products with their rows interleaved in program order: four carry chains in
// flight instead of two.  One warp alone runs a product as a ~500-cycle dependent chain (the assembler
// keeps the `volatile` carry-chain statements in order); latency-bound code with two products ready
// (the prefix and suffix scans of fused.cuh) overlaps them this way.
template <int N>
ECB_DEV void mul_full2(u32* t1, const u32* a1, const u32* b1, u32* t2, const u32* a2, const u32* b2) {
    u32 E1[2 * N + 2], O1[2 * N + 2], E2[2 * N + 2], O2[2 * N + 2];
    ECB_UNROLL
    for (int i = 0; i < 2 * N + 2; i++) { E1[i] = 0; O1[i] = 0; E2[i] = 0; O2[i] = 0; }
    ECB_UNROLL
    for (int i = 0; i < N; i++) {
        if ((i & 1) == 0) {
            mac_chain<N / 2, true>(E1 + i, a1, b1[i]);
            mac_chain<N / 2, true>(E2 + i, a2, b2[i]);
            mac_chain<N / 2, true>(O1 + i + 1, a1 + 1, b1[i]);
            mac_chain<N / 2, true>(O2 + i + 1, a2 + 1, b2[i]);
        } else {
            mac_chain<N / 2, true>(E1 + i + 1, a1 + 1, b1[i]);
            mac_chain<N / 2, true>(E2 + i + 1, a2 + 1, b2[i]);
            mac_chain<N / 2, true>(O1 + i, a1, b1[i]);
            mac_chain<N / 2, true>(O2 + i, a2, b2[i]);
        }
    }
    t1[0] = E1[0];
    t1[1] = add_cc(E1[1], O1[1]);
    ECB_UNROLL
    for (int i = 2; i < 2 * N; i++) t1[i] = addc_cc(E1[i], O1[i]);
    (void)addc(0, 0);
    t2[0] = E2[0];
    t2[1] = add_cc(E2[1], O2[1]);
    ECB_UNROLL
    for (int i = 2; i < 2 * N; i++) t2[i] = addc_cc(E2[i], O2[i]);
    (void)addc(0, 0);
}

// Full square t[0..2N) = a*a: off-diagonal products once, doubled, plus the
// diagonal.  N(N-1)/2 + N IMAD.WIDE.
template <int N>
ECB_DEV void sqr_full(u32* t, const u32* a) {
    u32 E[2 * N + 2], O[2 * N + 2];
    ECB_UNROLL
    for (int i = 0; i < 2 * N + 2; i++) { E[i] = 0; O[i] = 0; }
    // off-diagonal: for each i, products a_j*a_i with j>i.
    ECB_UNROLL
    for (int i = 0; i < N - 1; i++) {
        // j = i+1, i+3, ...  (i+j odd) -> O at limb 2i+1
        {
            const int cnt = (N - 1 - i + 1) / 2;  // number of j in {i+1,i+3,..} < N
            if (cnt > 0) {
                u32* acc = O + 2 * i + 1;
                const u32* aa = a + i + 1;
                acc[0] = mad_lo_cc(aa[0], a[i], acc[0]);
                acc[1] = madc_hi_cc(aa[0], a[i], acc[1]);
                ECB_UNROLL
                for (int k = 1; k < cnt; k++) {
                    acc[2 * k] = madc_lo_cc(aa[2 * k], a[i], acc[2 * k]);
                    acc[2 * k + 1] = madc_hi_cc(aa[2 * k], a[i], acc[2 * k + 1]);
                }
                acc[2 * cnt] = addc(acc[2 * cnt], 0);
            }
        }
        // j = i+2, i+4, ... (i+j even) -> E at limb 2i+2
        {
            const int cnt = (N - 1 - i) / 2;
            if (cnt > 0) {
                u32* acc = E + 2 * i + 2;
                const u32* aa = a + i + 2;
                acc[0] = mad_lo_cc(aa[0], a[i], acc[0]);
                acc[1] = madc_hi_cc(aa[0], a[i], acc[1]);
                ECB_UNROLL
                for (int k = 1; k < cnt; k++) {
                    acc[2 * k] = madc_lo_cc(aa[2 * k], a[i], acc[2 * k]);
                    acc[2 * k + 1] = madc_hi_cc(aa[2 * k], a[i], acc[2 * k + 1]);
                }
                acc[2 * cnt] = addc(acc[2 * cnt], 0);
            }
        }
    }
    // S = E + O  (limbs 1 .. 2N-1; limb 0 of the off-diagonal sum is 0)
    u32 S[2 * N];
    S[0] = 0;
    S[1] = add_cc(E[1], O[1]);
    ECB_UNROLL
    for (int i = 2; i < 2 * N; i++) S[i] = addc_cc(E[i], O[i]);
    // S *= 2
    S[1] = add_cc(S[1], S[1]);
    ECB_UNROLL
    for (int i = 2; i < 2 * N; i++) S[i] = addc_cc(S[i], S[i]);
    // diagonal a_i^2 at limbs (2i, 2i+1), added on the fly
    {
        u32 lo = mul_lo(a[0], a[0]), hi = mul_hi(a[0], a[0]);
        t[0] = lo;
        t[1] = add_cc(S[1], hi);
    }
    ECB_UNROLL
    for (int i = 1; i < N; i++) {
        u32 lo = mul_lo(a[i], a[i]), hi = mul_hi(a[i], a[i]);
        t[2 * i] = addc_cc(S[2 * i], lo);
        t[2 * i + 1] = addc_cc(S[2 * i + 1], hi);
    }
    (void)addc(0, 0);
}

}  // namespace ecb
