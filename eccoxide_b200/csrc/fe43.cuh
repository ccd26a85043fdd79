// fe43.cuh — GF(2^255-19) on the FP64 pipe: 6 signed limbs of 43 bits held in doubles.
//
// The integer field (fe25519.cuh) keeps the integer-multiply pipe 75-86 % busy in every kernel while
// the FP64 pipe idles; on B200 DFMA issues at 64 / clk / SM (measured: tu_probe.cu variant 4), twice
// the IMAD.WIDE rate.  This is the same field on that other pipe, so that half the warps of a kernel
// can run on each (verdict r01 item 3; the reference's own radix-2^51 layout,
// src/curve/fiat/curve25519_64.rs:217, does not survive double rounding: five 51-bit limbs leave
// no headroom in a 53-bit significand, six 43-bit limbs do).
//
// value = sum v[i] * 2^(43 i), every v[i] an integer-valued double, SIGNED.  Exactness argument:
//   product split   h = RN(a b / 2^43) 2^43 is read off an FMA against 1.5 * 2^95 (ulp 2^43), the
//                   running sum of a column's h's lives in that same FMA chain; l = fma(a, b, -h) is the
//                   exact remainder, |l| <= 2^42.  Needs |sum of a column's a_i b_j| < 2^94:
//                   |limbs| <= 2^45.5 on input.
//   carries         c = RN(V / 2^43) by the 1.5 * 2^52 trick, V - c 2^43 by one FMA: all values stay
//                   below 2^52 in magnitude, so every operation is exact integer arithmetic.
//   2^258 = 152 (mod p) folds columns 6..11 into 0..5.
// Output limbs of mul / sqr are "tight": |v[i]| <= 2^42 + 2^16.  add / sub are six DADDs, no carry.
// Plain C++ (fma from <cmath> is correctly rounded): the same code runs in tests/hostsim.
#pragma once
#include "fe25519.cuh"
#ifdef ECB_HOSTSIM
#include <cmath>
#endif

namespace ecb {

struct fe43 {
    double v[6];
};

struct F43 {
    typedef fe43 el;
    static constexpr double M52 = 6755399441055744.0;                       // 1.5 * 2^52
    static constexpr double T43 = 8796093022208.0;                          // 2^43
    static constexpr double R43 = 1.0 / 8796093022208.0;                    // 2^-43
    static constexpr double M95 = 6755399441055744.0 * 8796093022208.0;     // 1.5 * 2^95

    ECB_DEV static double fma_(double a, double b, double c) {
#ifdef ECB_HOSTSIM
        return std::fma(a, b, c);
#else
        return __fma_rn(a, b, c);
#endif
    }
    ECB_DEV static void add(el& r, const el& a, const el& b) {
        ECB_UNROLL
        for (int i = 0; i < 6; i++) r.v[i] = a.v[i] + b.v[i];
    }
    ECB_DEV static void sub(el& r, const el& a, const el& b) {
        ECB_UNROLL
        for (int i = 0; i < 6; i++) r.v[i] = a.v[i] - b.v[i];
    }
    ECB_DEV static void dbl(el& r, const el& a) {
        ECB_UNROLL
        for (int i = 0; i < 6; i++) r.v[i] = a.v[i] + a.v[i];
    }
    ECB_DEV static void neg(el& r, const el& a) {
        ECB_UNROLL
        for (int i = 0; i < 6; i++) r.v[i] = -a.v[i];
    }
    // columns H (running sums of the rounded high parts, biased by M95) and L (exact low parts) -> 6 tight limbs
    ECB_DEV static void reduce(el& r, const double* H, const double* L) {
        double V[13];
        V[0] = L[0];
        ECB_UNROLL
        for (int k = 1; k < 11; k++) V[k] = fma_(H[k - 1] - M95, R43, L[k]);
        V[11] = (H[10] - M95) * R43;
        // pass 1: columns 5..11, all carries at once
        double c[12];
        ECB_UNROLL
        for (int k = 5; k < 12; k++) {
            c[k] = fma_(V[k], R43, M52) - M52;
            V[k] = fma_(c[k], -T43, V[k]);
        }
        ECB_UNROLL
        for (int k = 6; k < 12; k++) V[k] += c[k - 1];
        V[12] = c[11];
        // fold: 2^(43 (k + 6)) = 152 * 2^(43 k), 2^(43 * 12) = 152^2
        double t[6];
        ECB_UNROLL
        for (int k = 0; k < 6; k++) t[k] = fma_(V[k + 6], 152.0, V[k]);
        t[0] = fma_(V[12], 23104.0, t[0]);
        // pass 2
        double d[6];
        ECB_UNROLL
        for (int k = 0; k < 6; k++) {
            d[k] = fma_(t[k], R43, M52) - M52;
            t[k] = fma_(d[k], -T43, t[k]);
        }
        r.v[0] = fma_(d[5], 152.0, t[0]);
        ECB_UNROLL
        for (int k = 1; k < 6; k++) r.v[k] = t[k] + d[k - 1];
    }
    ECB_DEV static void mul(el& r, const el& a, const el& b) {
        double H[11], L[11];
        ECB_UNROLL
        for (int k = 0; k < 11; k++) { H[k] = M95; L[k] = 0.0; }
        ECB_UNROLL
        for (int i = 0; i < 6; i++) {
            ECB_UNROLL
            for (int j = 0; j < 6; j++) {
                double hn = fma_(a.v[i], b.v[j], H[i + j]);
                double h = hn - H[i + j];
                H[i + j] = hn;
                L[i + j] += fma_(a.v[i], b.v[j], -h);
            }
        }
        reduce(r, H, L);
    }
    ECB_DEV static void sqr(el& r, const el& a) {
        double H[11], L[11], a2[6];
        ECB_UNROLL
        for (int k = 0; k < 11; k++) { H[k] = M95; L[k] = 0.0; }
        ECB_UNROLL
        for (int i = 0; i < 6; i++) a2[i] = a.v[i] + a.v[i];
        ECB_UNROLL
        for (int i = 0; i < 6; i++) {
            ECB_UNROLL
            for (int j = i; j < 6; j++) {
                double x = a.v[i], y = (j == i) ? a.v[j] : a2[j];
                double hn = fma_(x, y, H[i + j]);
                double h = hn - H[i + j];
                H[i + j] = hn;
                L[i + j] += fma_(x, y, -h);
            }
        }
        reduce(r, H, L);
    }

    // ---- conversions --------------------------------------------------------------------------
    // 8 little-endian words (any 256-bit value) -> limbs in [0, 2^43) (limb 5: < 2^41)
    ECB_DEV static void from_words(el& r, const u32* w) {
        ECB_UNROLL
        for (int i = 0; i < 6; i++) {
            const int bit = 43 * i, wd = bit >> 5, sh = bit & 31;
            // 43 bits starting at (wd, sh): up to three words
            u64 x = (u64)w[wd] >> sh;
            if (wd + 1 < 8) x |= (u64)w[wd + 1] << (32 - sh);
            if (wd + 2 < 8 && sh > 21) x |= (u64)w[wd + 2] << (64 - sh);
            x &= ((u64)1 << 43) - 1;
            r.v[i] = u64_to_double(x);
        }
    }
    ECB_DEV static double u64_to_double(u64 x) {   // x < 2^52
#ifdef ECB_HOSTSIM
        return (double)x;
#else
        return __longlong_as_double((long long)(x | 0x4330000000000000ull)) - 4503599627370496.0;
#endif
    }
    // limbs (|v[i]| <= 2^45.5) -> 8 words of the integer field, value congruent mod p ("loose" fe25519)
    ECB_DEV static void to_fe25519(fe25519& r, const el& a) {
        // bias every limb by 2^46 so that it is positive, subtract the bias value afterwards:
        // B = 2^46 * sum 2^(43 i) mod p is a constant of the integer field
        u32 acc[10];
        ECB_UNROLL
        for (int i = 0; i < 10; i++) acc[i] = 0;
        ECB_UNROLL
        for (int i = 0; i < 6; i++) {
            double x = a.v[i] + 70368744177664.0;                 // + 2^46: in (0, 2^47)
#ifdef ECB_HOSTSIM
            u64 u = (u64)x;
#else
            u64 u = (u64)__double_as_longlong(x + 4503599627370496.0) & 0x000fffffffffffffull;
#endif
            const int bit = 43 * i, wd = bit >> 5, sh = bit & 31;
            // add u << sh at word wd (u < 2^47, sh < 32: three words)
            u64 lo = u << sh;
            u32 hi = sh ? (u32)(u >> (64 - sh)) : 0u;
            acc[wd] = add_cc(acc[wd], (u32)lo);
            acc[wd + 1] = addc_cc(acc[wd + 1], (u32)(lo >> 32));
            acc[wd + 2] = addc_cc(acc[wd + 2], hi);
            ECB_UNROLL
            for (int k = wd + 3; k < 10; k++) acc[k] = addc_cc(acc[k], 0);
            (void)addc(0, 0);
        }
        // acc < 2^(43*5 + 47 + 1) = 2^263: fold words 8, 9 (2^256 = 38)
        fe25519 t;
        u32 R[10];
        ECB_UNROLL
        for (int i = 0; i < 8; i++) R[i] = acc[i];
        R[8] = 0;
        R[9] = 0;
        R[0] = mad_lo_cc(acc[8], 38u, R[0]);
        R[1] = madc_hi_cc(acc[8], 38u, R[1]);
        ECB_UNROLL
        for (int i = 2; i < 8; i++) R[i] = addc_cc(R[i], 0);
        u32 c = addc(0, 0);
        R[0] = add_cc(R[0], 38u * c);
        ECB_UNROLL
        for (int i = 1; i < 8; i++) R[i] = addc_cc(R[i], 0);
        c = addc(0, 0);
        R[0] += 38u * c;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) t.v[i] = R[i];
        (void)acc[9];   // 2^263 bound: acc[9] is always 0
        // bias = 2^46 (1 + 2^43 + 2^86 + 2^129 + 2^172 + 2^215) mod p = 2^46 + 2^89 + 2^132 + 2^175 + 2^218 + 64 * 19
        fe25519 bias;
        bias.v[0] = 0x000004c0u; bias.v[1] = 0x00004000u; bias.v[2] = 0x02000000u; bias.v[3] = 0u;
        bias.v[4] = 0x00000010u; bias.v[5] = 0x00008000u; bias.v[6] = 0x04000000u; bias.v[7] = 0u;
        F25519::sub(r, t, bias);
    }
};

}  // namespace ecb
