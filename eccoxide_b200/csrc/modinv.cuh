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
#include "params_gen.cuh"

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


// The same 30 divsteps, four at a time through a table.  Four steps depend only on the low four bits of f (odd)
// and g and on where zeta sits relative to zero, so their combined transition matrix and the new zeta are
// tabulated (SG_JUMP4, 1408 words, generated and documented by tools/gen_params.py: safegcd_jump_table): seven
// lookups and two single steps replace ~10 data-dependent trips of sg_divsteps30_var.  With one warp per SM —
// the block-level inversions — a trip of that loop costs ~120 cycles (BREV + FLO for the ctz, a dependent chain
// of ~25 instructions); a lookup is one shared-memory load plus ~12 multiply-adds that do not wait on each other.
// tbl: the table in shared memory (device) or SG_JUMP4 itself (host simulation).
ECB_DEV s32 sg_divsteps30_jump(s32 zeta, u32 f0, u32 g0, s32& tu, s32& tv, s32& tq, s32& tr, const u32* tbl) {
    u32 U = 1, V = 0, Q = 0, R = 1, f = f0, g = g0;
    ECB_UNROLL
    for (int it = 0; it < 7; it++) {
        s32 zc = zeta < -6 ? -6 : (zeta > 4 ? 4 : zeta);
        const u32 e = tbl[((u32)(zc + 6) * 8u + ((f >> 1) & 7u)) * 16u + (g & 15u)];
        const u32 mu = (u32)((s32)(e << 26) >> 26), mv = (u32)((s32)(e << 20) >> 26);
        const u32 mq = (u32)((s32)(e << 14) >> 26), mr = (u32)((s32)(e << 8) >> 26);
        const s32 off = (s32)(e << 2) >> 26;
        zeta = ((e >> 30) & 1u ? -zeta : zeta) + off;
        const u32 nf = (mu * f + mv * g) >> 4, ng = (mq * f + mr * g) >> 4;
        const u32 nU = mu * U + mv * Q, nV = mu * V + mv * R, nQ = mq * U + mr * Q, nR = mq * V + mr * R;
        f = nf; g = ng; U = nU; V = nV; Q = nQ; R = nR;
    }
    ECB_UNROLL
    for (int i = 0; i < 2; i++) {   // steps 29 and 30, one at a time (the loop body of sg_divsteps30)
        u32 c1 = (u32)(zeta >> 31);
        u32 c2 = 0u - (g & 1u);
        u32 x = (f ^ c1) - c1, y = (U ^ c1) - c1, z = (V ^ c1) - c1;
        g += x & c2;
        Q += y & c2;
        R += z & c2;
        c1 &= c2;
        zeta = (zeta ^ (s32)c1) - 1;
        f += g & c1;
        U += Q & c1;
        V += R & c1;
        g >>= 1;
        U <<= 1;
        V <<= 1;
    }
    tu = (s32)U; tv = (s32)V; tq = (s32)Q; tr = (s32)R;
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

#ifndef ECB_HOSTSIM
// ---------------------------------------------------------------------------------------------------
// The same inverse computed by ONE WARP for ONE element (every lane passes the same a and receives the
// same result).  The batch-inversion kernels run a single inversion per block while every other warp
// waits, so its latency — not its instruction count — is what a small batch pays.  In sg_modinv one
// thread spends about half of each batch of 30 divsteps in four serial NL-limb carry chains.  Here
//   lanes 0 .. NL-1        hold limb j of (f, g),
//   lanes 16 .. 16+NL-1    hold limb j of (d, e)          (NL <= 16),
// every lane derives the transition matrix from the broadcast low limbs (identical work, no divergence),
// multiplies its own limb pair, and the carries move ONE lane per batch: limbs stay signed and only
// partially normalised (in [-8, 2^30 + 8]; the low limb is exact mod 2^30, which is all the divsteps
// read).  The sign of d, e is then read from the top limb alone; a wrong guess within 2^-40 of zero
// loosens the (-2p, p) bound by one p, which the final reduction absorbs.  Ranges and results are
// checked lane for lane in tools/models/safegcd_warp_model.py.  The divsteps come from the jump table
// (sg_divsteps30_jump), which the caller stages in shared memory with sg_stage_jump_table.
// ---------------------------------------------------------------------------------------------------
// every thread of the block copies its share of SG_JUMP4 into `sh` (SG_JUMP_WORDS words); the caller synchronises
__device__ __forceinline__ void sg_stage_jump_table(u32* sh) {
    for (int i = (int)threadIdx.x; i < SG_JUMP_WORDS; i += (int)blockDim.x) sh[i] = SG_JUMP4[i];
}
template <int NW, int NL, int MAXB>
__device__ __forceinline__ void sg_modinv_warp(u32* r, const u32* a, const u32* p, const u32* jump) {
    static_assert(NL <= 16, "one limb per lane: NL <= 16");
    const unsigned FULL = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31u), j = lane & 15;
    const bool de = lane >= 16, live = j < NL, top = j == NL - 1;
    const s32 M30 = 0x3fffffff;
    s32 ml[NL], al[NL];
    sg_from_words<NW, NL>(ml, p);
    sg_from_words<NW, NL>(al, a);
    s32 mj = 0, x = 0, y = 0;   // this lane's limb of the modulus, of f | d, of g | e
    ECB_UNROLL
    for (int i = 0; i < NL; i++) {
        if (j == i) { mj = ml[i]; x = de ? 0 : ml[i]; y = de ? (i == 0 ? 1 : 0) : al[i]; }
    }
    if (!de) mj = 0;            // the multiple of p only enters (d, e)
    u32 pinv = (u32)ml[0];
    ECB_UNROLL
    for (int i = 0; i < 5; i++) pinv *= 2u - (u32)ml[0] * pinv;
    pinv &= 0x3fffffffu;
    s32 zeta = -1;
    ECB_NOUNROLL
    for (int b = 0; b < MAXB; b++) {
        const u32 f0 = (u32)__shfl_sync(FULL, x, 0), g0 = (u32)__shfl_sync(FULL, y, 0);
        const s32 d0 = __shfl_sync(FULL, x, 16), e0 = __shfl_sync(FULL, y, 16);
        const s32 sd = __shfl_sync(FULL, x, 16 + NL - 1) >> 31, se = __shfl_sync(FULL, y, 16 + NL - 1) >> 31;
        s32 u, v, q, rr;
        zeta = sg_divsteps30_jump(zeta, f0, g0, u, v, q, rr, jump);   // jump: SG_JUMP4 staged in shared memory
        s32 md = (u & sd) + (v & se), me = (q & sd) + (rr & se);
        const u32 cd = (u32)u * (u32)d0 + (u32)v * (u32)e0, ce = (u32)q * (u32)d0 + (u32)rr * (u32)e0;
        md -= (s32)((pinv * cd + (u32)md) & (u32)M30);
        me -= (s32)((pinv * ce + (u32)me) & (u32)M30);
        const s64 tx = (s64)u * x + (s64)v * y + (s64)mj * md;
        const s64 ty = (s64)q * x + (s64)rr * y + (s64)mj * me;
        // divide by 2^30: limb j of the quotient = high part of product j + low 30 bits of product j + 1
        s32 lox = __shfl_down_sync(FULL, (s32)tx & M30, 1), loy = __shfl_down_sync(FULL, (s32)ty & M30, 1);
        if (top || !live) { lox = 0; loy = 0; }
        const s64 nx = (tx >> 30) + lox, ny = (ty >> 30) + loy;
        s32 cx = top ? 0 : (s32)(nx >> 30), cy = top ? 0 : (s32)(ny >> 30);
        x = top ? (s32)nx : ((s32)nx & M30);
        y = top ? (s32)ny : ((s32)ny & M30);
        cx = __shfl_up_sync(FULL, cx, 1);
        cy = __shfl_up_sync(FULL, cy, 1);
        if (j > 0 && live) { x += cx; y += cy; }
        if (!live) { x = 0; y = 0; }
        if (__ballot_sync(FULL, !de && y != 0) == 0u) break;   // every limb of g is zero: g = 0
    }
    // f = +-1 (its low limb is exact: 1 or 2^30 - 1); d = +-a^-1 somewhere in (-3p, 2p), limbs loose
    const s32 fneg = ((u32)__shfl_sync(FULL, x, 0) == 1u) ? 0 : -1;
    s32 d[NL];
    ECB_UNROLL
    for (int i = 0; i < NL; i++) d[i] = __shfl_sync(FULL, x, 16 + i);
    ECB_UNROLL
    for (int i = 0; i < NL; i++) d[i] = (d[i] ^ fneg) - fneg;
    ECB_UNROLL
    for (int rep = 0; rep < 6; rep++) {
        // rep 0: carry only; rep 1-3: add p while negative; rep 4-5: subtract p while >= p (tested as d - p >= 0)
        s32 t[NL], c = 0;
        const s32 neg = d[NL - 1] >> 31;
        ECB_UNROLL
        for (int i = 0; i < NL; i++) {
            s32 w = d[i] + c;
            if (rep >= 1 && rep <= 3) w += ml[i] & neg;
            if (rep >= 4) w -= ml[i];
            if (i < NL - 1) { t[i] = w & M30; c = w >> 30; } else t[i] = w;
        }
        const s32 keep = (rep >= 4) ? ~(t[NL - 1] >> 31) : -1;   // a subtraction that went negative is dropped
        ECB_UNROLL
        for (int i = 0; i < NL; i++) d[i] = (t[i] & keep) | (d[i] & ~keep);
    }
    sg_to_words<NW, NL>(r, d);
}
#endif

}  // namespace ecb
