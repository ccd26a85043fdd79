// modinv.cuh — modular inversion by the Bernstein–Yang "safegcd" divsteps on signed 30-bit limbs.
//
// Replaces the Fermat exponentiations of the batch-inversion kernels (one inversion per thread:
// a^(p-2) = 254 S + 11 M for 2^255-19, a 265-step dependent chain of field operations whose latency
// was the floor of k_batch_inv) with ~600 (256-bit) / ~900 (384-bit) divsteps processed 30 at a time:
// per batch a 2x2 transition matrix from the low words of f and g (simple alu instructions), then
// f, g, d, e are updated with 10 N30 small multiplications.  About 6x fewer passes through the
// integer-multiply pipe and a 6x shorter dependent chain than the Fermat chain.
//
// The reference uses the same family of algorithm for its generic inverse (inverse_safegcd,
// src/curve/fiat/field_macros.rs:692-770, used by bls12_381 Fp, fp.rs:55-57) and Fermat chains for
// curve25519 / p256r1 / p384r1 (curve25519.rs:155-200, p256r1.rs:49-65, p384r1.rs:50-69); the value
// of a modular inverse does not depend on how it is computed.
//
// Variable time (the loop stops when g = 0): fine for the public-data batch path.
// Plain integer C++: the same code runs in the host simulation.
#pragma once
#include "limb.cuh"

namespace ecb {

typedef int32_t s32;
typedef int64_t s64;

// 30 divsteps on the low words of f, g (f odd); returns the new zeta and the transition matrix
// t = (u, v; q, r) with [f', g'] = t [f, g] / 2^30.
ECB_DEV s32 sg_divsteps30(s32 zeta, u32 f0, u32 g0, s32& tu, s32& tv, s32& tq, s32& tr) {
    u32 u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
    ECB_NOUNROLL
    for (int i = 0; i < 30; i++) {
        u32 c1 = (u32)(zeta >> 31);
        u32 c2 = 0u - (g & 1u);
        u32 x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
        g += x & c2;
        q += y & c2;
        r += z & c2;
        c1 &= c2;
        zeta = (zeta ^ (s32)c1) - 1;
        f += g & c1;
        u += q & c1;
        v += r & c1;
        g >>= 1;
        u <<= 1;
        v <<= 1;
    }
    tu = (s32)u; tv = (s32)v; tq = (s32)q; tr = (s32)r;
    return zeta;
}

// The same 30 divsteps in variable time: runs of zero low bits of g are shifted out at once (ctz), and
// while no swap can occur (zeta >= 0 for the next `limit` steps) up to four low bits of g are cancelled
// with one multiple of f, w = -g / f mod 2^limit (f (f^2 - 2) = -1/f mod 16 for odd f).  About 10 loop
// iterations per batch instead of 30 (measured over random inputs: 184 iterations in 17.9 batches for a
// 255-bit inverse), each a short chain of alu instructions — the batch-inversion kernels run ONE such
// chain per block while every other warp waits, so its latency is the floor of a small batch.
// Same transition matrix, same zeta as 30 calls of the single step: [f', g'] = t [f, g] / 2^30.
ECB_DEV int sg_ctz(u32 x) {   // x != 0
#ifdef ECB_HOSTSIM
    return __builtin_ctz(x);
#else
    return __ffs((int)x) - 1;
#endif
}
ECB_DEV s32 sg_divsteps30_var(s32 zeta, u32 f0, u32 g0, s32& tu, s32& tv, s32& tq, s32& tr) {
    u32 u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
    int i = 30;
    for (;;) {
        int z = sg_ctz(g | (1u << i));   // at most i
        g >>= z;
        u <<= z;
        v <<= z;
        zeta -= z;
        i -= z;
        if (i == 0) break;
        if (zeta < 0) {                  // g odd, delta > 0: (f, g) <- (g, -f), the step itself follows below
            zeta = -zeta - 1;
            u32 t = f; f = g; g = 0u - t;
            t = u; u = q; q = 0u - t;
            t = v; v = r; r = 0u - t;
        }
        int limit = zeta + 1 < i ? zeta + 1 : i;   // steps that cannot swap
        if (limit > 4) limit = 4;
        u32 m = (1u << limit) - 1u;
        u32 w = (f * g * (f * f - 2u)) & m;         // -g / f mod 2^limit
        g += f * w;
        q += u * w;
        r += v * w;
    }
    tu = (s32)u; tv = (s32)v; tq = (s32)q; tr = (s32)r;
    return zeta;
}

// NW 32-bit words (little-endian, value < 2^(32 NW)) <-> NL signed 30-bit limbs (30 NL >= 32 NW + 2)
template <int NW, int NL>
ECB_DEV void sg_from_words(s32* l, const u32* w) {
    ECB_UNROLL
    for (int i = 0; i < NL; i++) {
        int bit = 30 * i, wd = bit >> 5, sh = bit & 31;
        u32 lo = wd < NW ? w[wd] : 0u, hi = (wd + 1) < NW ? w[wd + 1] : 0u;
        u32 v = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
        l[i] = (s32)(v & 0x3fffffffu);
    }
}
template <int NW, int NL>
ECB_DEV void sg_to_words(u32* w, const s32* l) {  // limbs in [0, 2^30)
    ECB_UNROLL
    for (int i = 0; i < NW; i++) {
        int bit = 32 * i, li = bit / 30, sh = bit % 30;
        u32 v = (u32)l[li] >> sh;
        if (li + 1 < NL) v |= (u32)l[li + 1] << (30 - sh);
        if (sh > 28 && li + 2 < NL) v |= (u32)l[li + 2] << (60 - sh);
        w[i] = v;
    }
}

// r = a^-1 mod p for 0 < a < p given as NW words; p odd.  a = 0 gives 0 (invert_or_zero).
// MAXB: batches of 30 divsteps; 590 steps suffice for 256-bit, 886 for 384-bit, 1033 for 448-bit moduli.
template <int NW, int NL, int MAXB>
ECB_DEV void sg_modinv(u32* r, const u32* a, const u32* p) {
    s32 f[NL], g[NL], d[NL], e[NL], m[NL];
    sg_from_words<NW, NL>(m, p);
    sg_from_words<NW, NL>(g, a);
    u32 az = 0;
    ECB_UNROLL
    for (int i = 0; i < NL; i++) { f[i] = m[i]; d[i] = 0; e[i] = 0; az |= (u32)g[i]; }
    e[0] = 1;
    if (az == 0) {
        ECB_UNROLL
        for (int i = 0; i < NW; i++) r[i] = 0;
        return;
    }
    // p^-1 mod 2^30 (Newton; p odd)
    u32 pinv = (u32)m[0];
    ECB_UNROLL
    for (int i = 0; i < 5; i++) pinv *= 2u - (u32)m[0] * pinv;
    pinv &= 0x3fffffffu;
    const s32 M30 = 0x3fffffff;
    s32 zeta = -1;
    ECB_NOUNROLL
    for (int b = 0; b < MAXB; b++) {
        s32 u, v, q, rr;
        zeta = sg_divsteps30_var(zeta, (u32)f[0], (u32)g[0], u, v, q, rr);
        {   // (d, e) <- t (d, e) / 2^30 mod p, kept in (-2p, p)
            s32 sd = d[NL - 1] >> 31, se = e[NL - 1] >> 31;
            s32 md = (u & sd) + (v & se), me = (q & sd) + (rr & se);
            s64 cd = (s64)u * d[0] + (s64)v * e[0], ce = (s64)q * d[0] + (s64)rr * e[0];
            md -= (s32)((pinv * (u32)cd + (u32)md) & (u32)M30);
            me -= (s32)((pinv * (u32)ce + (u32)me) & (u32)M30);
            cd += (s64)m[0] * md;
            ce += (s64)m[0] * me;
            cd >>= 30;
            ce >>= 30;
            ECB_UNROLL
            for (int i = 1; i < NL; i++) {
                cd += (s64)u * d[i] + (s64)v * e[i] + (s64)m[i] * md;
                ce += (s64)q * d[i] + (s64)rr * e[i] + (s64)m[i] * me;
                d[i - 1] = (s32)cd & M30;
                cd >>= 30;
                e[i - 1] = (s32)ce & M30;
                ce >>= 30;
            }
            d[NL - 1] = (s32)cd;
            e[NL - 1] = (s32)ce;
        }
        {   // (f, g) <- t (f, g) / 2^30
            s64 cf = (s64)u * f[0] + (s64)v * g[0], cg = (s64)q * f[0] + (s64)rr * g[0];
            cf >>= 30;
            cg >>= 30;
            u32 gz = 0;
            ECB_UNROLL
            for (int i = 1; i < NL; i++) {
                cf += (s64)u * f[i] + (s64)v * g[i];
                cg += (s64)q * f[i] + (s64)rr * g[i];
                f[i - 1] = (s32)cf & M30;
                cf >>= 30;
                g[i - 1] = (s32)cg & M30;
                cg >>= 30;
                gz |= (u32)g[i - 1];
            }
            f[NL - 1] = (s32)cf;
            g[NL - 1] = (s32)cg;
            gz |= (u32)g[NL - 1];
            if (gz == 0) break;
        }
    }
    // f = +-1, d in (-2p, p): result = sign(f) d mod p
    s32 fneg = f[NL - 1] >> 31;
    ECB_UNROLL
    for (int rep = 0; rep < 3; rep++) {
        // rep 0: if d < 0 add p; rep 1: negate if f < 0 (then value in (-p, p)); rep 2: if < 0 add p
        s32 add = d[NL - 1] >> 31;
        if (rep == 1) {
            ECB_UNROLL
            for (int i = 0; i < NL; i++) d[i] = (d[i] ^ fneg) - fneg;
            add = 0;
        }
        s32 c = 0;
        ECB_UNROLL
        for (int i = 0; i < NL; i++) {
            s32 t = d[i] + (m[i] & add) + c;
            if (i < NL - 1) { d[i] = t & M30; c = t >> 30; } else d[i] = t;
        }
    }
    sg_to_words<NW, NL>(r, d);
}

}  // namespace ecb
