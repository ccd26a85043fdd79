// kernels.cuh — per-thread bodies of the batch kernels.
//
// Each body takes its global thread index explicitly; eccbatch.cu wraps them in __global__
// launchers (one thread per scalar multiplication), tests/hostsim/ loops over them on the CPU.
//
// Data layout in HBM
//   inputs / outputs : the reference's wire encodings, AoS, contiguous (32/48/56-byte elements)
//   intermediates    : structure-of-arrays "planes": plane[limb * n + idx], so a warp reading limb
//                      `l` of 32 consecutive elements touches one 128-byte line
//   tables           : fixed-base comb tables as arrays of niels entries (96 B, packed or on a 128-byte
//                      stride) in HBM: 0.65-43 GB depending on the window width
//                      per-thread window tables of variable-base kernels in a scratch arena,
//                      one contiguous block per resident thread
#pragma once
#include "edwards.cuh"
#include "weier.cuh"

namespace ecb {


#ifdef ECB_HOSTSIM
ECB_DEV void report_bad(unsigned long long* st, size_t idx, u32 code) {
    unsigned long long v = ((unsigned long long)idx << 8) | code;
    if (v < *st) *st = v;
}
ECB_DEV u32 bswap32(u32 x) { return __builtin_bswap32(x); }
#else
ECB_DEV void report_bad(unsigned long long* st, size_t idx, u32 code) {
    atomicMin(st, ((unsigned long long)idx << 8) | code);
}
ECB_DEV u32 bswap32(u32 x) { return __byte_perm(x, 0, 0x0123); }
#endif

// ---- word movers -----------------------------------------------------------------------
template <int NW>
ECB_DEV void ld_words(u32* dst, const u32* src) {
#ifndef ECB_HOSTSIM
    if (NW % 4 == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        ECB_UNROLL
        for (int i = 0; i < NW / 4; i++) {
            uint4 v = __ldg(s4 + i);
            dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
        }
        return;
    }
    if (NW % 2 == 0) {
        const uint2* s2 = reinterpret_cast<const uint2*>(src);
        ECB_UNROLL
        for (int i = 0; i < NW / 2; i++) {
            uint2 v = __ldg(s2 + i);
            dst[2 * i] = v.x; dst[2 * i + 1] = v.y;
        }
        return;
    }
#endif
    ECB_UNROLL
    for (int i = 0; i < NW; i++) dst[i] = src[i];
}
// coherent variant for memory written earlier by the same kernel (per-thread scratch tables):
// ld.global.nc (__ldg) must not be used there.
template <int NW>
ECB_DEV void ld_words_rw(u32* dst, const u32* src) {
#ifndef ECB_HOSTSIM
    if (NW % 4 == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        ECB_UNROLL
        for (int i = 0; i < NW / 4; i++) {
            uint4 v = s4[i];
            dst[4 * i] = v.x; dst[4 * i + 1] = v.y; dst[4 * i + 2] = v.z; dst[4 * i + 3] = v.w;
        }
        return;
    }
#endif
    ECB_UNROLL
    for (int i = 0; i < NW; i++) dst[i] = src[i];
}
template <int NW>
ECB_DEV void st_words(u32* dst, const u32* src) {
#ifndef ECB_HOSTSIM
    if (NW % 4 == 0) {
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        ECB_UNROLL
        for (int i = 0; i < NW / 4; i++) d4[i] = make_uint4(src[4 * i], src[4 * i + 1], src[4 * i + 2], src[4 * i + 3]);
        return;
    }
    if (NW % 2 == 0) {
        uint2* d2 = reinterpret_cast<uint2*>(dst);
        ECB_UNROLL
        for (int i = 0; i < NW / 2; i++) d2[i] = make_uint2(src[2 * i], src[2 * i + 1]);
        return;
    }
#endif
    ECB_UNROLL
    for (int i = 0; i < NW; i++) dst[i] = src[i];
}
// big-endian wire bytes (NW words) <-> little-endian limbs
template <int NW>
ECB_DEV void ld_words_be(u32* dst, const u32* src) {
    u32 t[NW];
    ld_words<NW>(t, src);
    ECB_UNROLL
    for (int i = 0; i < NW; i++) dst[i] = bswap32(t[NW - 1 - i]);
}
template <int NW>
ECB_DEV void st_words_be(u32* dst, const u32* src) {
    u32 t[NW];
    ECB_UNROLL
    for (int i = 0; i < NW; i++) t[i] = bswap32(src[NW - 1 - i]);
    st_words<NW>(dst, t);
}
template <int NW>
ECB_DEV void plane_st(u32* plane, size_t n, size_t idx, const u32* v) {
    ECB_UNROLL
    for (int i = 0; i < NW; i++) plane[(size_t)i * n + idx] = v[i];
}
// hint: bring the NW lines of element idx (one per limb plane) towards the SM; no registers held
template <int NW>
ECB_DEV void plane_prefetch(const u32* plane, size_t n, size_t idx) {
#ifndef ECB_HOSTSIM
    ECB_UNROLL
    for (int i = 0; i < NW; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(plane + (size_t)i * n + idx));
#else
    (void)plane; (void)n; (void)idx;
#endif
}
template <int NW>
ECB_DEV void plane_ld(u32* v, const u32* plane, size_t n, size_t idx) {
    ECB_UNROLL
    for (int i = 0; i < NW; i++) v[i] = plane[(size_t)i * n + idx];
}

// l = 2^252 + 27742317777372353535851937790883648493 (curve25519.rs:46 ORDER_LIMBS)
ECB_CONST u32 ED25519_L[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu,
                              0x00000000u, 0x00000000u, 0x00000000u, 0x10000000u};
ECB_DEV u32 lt_words8(const u32* a, const u32* m) {  // a < m ?
    u32 d = sub_cc(a[0], m[0]);
    ECB_UNROLL
    for (int i = 1; i < 8; i++) d = subc_cc(a[i], m[i]);
    (void)d;
    return subc(0, 0) & 1;
}

// =======================================================================================
// Ed25519 fixed base: k*B from a signed-digit comb table  (replaces Point::mul_base,
// curve25519.rs:840-869 + params/comb/curve25519.rs)
//   table[(i * half + (j-1)) * 24 ..] = niels( j * 2^(W*i) * B ),  j = 1..half = 2^(W-1),
//   i = 0..nwin-1, nwin = ceil(254 / W).
// Writes projective X, Y, Z planes; the affine conversion is the batch-inversion kernel.
// =======================================================================================
// CLAMP: the scalar is an X25519 secret (protocol/x25519.rs:15 clamp, :49 x25519_base): any 32 bytes,
// clamped here, up to 255 bits — the comb covers W * nwin >= 256 bits and k*B = (k mod l)*B.
template <bool CLAMP>
ECB_DEV void ed25519_load_scalar(u32* k, size_t idx, const u32* scalars, unsigned long long* status) {
    ld_words<8>(k, scalars + idx * 8);
    k[8] = 0;
    if (CLAMP) {
        k[0] &= 0xfffffff8u;
        k[7] = (k[7] & 0x7fffffffu) | 0x40000000u;
    } else if (!lt_words8(k, ED25519_L)) {
        if (status) report_bad(status, idx, ST_NONCANONICAL_SCALAR);
        ECB_UNROLL
        for (int i = 0; i < 8; i++) k[i] = 0;
    }
}
// acc = sum of d_i * 2^(W i) * B over the windows i = first, first + step, ... < nwin (d_i = signed
// Booth digit of k).  The reference adds its 64 window entries in one chain (curve25519.rs:844-849);
// they are independent, so a scalar's windows can be split over `step` lanes and the partial sums
// added afterwards.  T of the result is computed only when want_t (a further addition follows).
// Entry (i, j) sits at table + (i * half + j - 1) * stride words (stride 24 packed, 32 = 128-byte aligned).
ECB_DEV void ed25519_comb_partial(ge_p3& acc, const u32* k, const u32* table, int W, int nwin, int stride, int first, int step,
                                  bool want_t) {
    const u32 half = 1u << (W - 1);
    ge_identity(acc);
    u32 v[8];                      // 2k as a shift register (k < 2^255, so 2k fits 256 bits)
    booth_reg_init<8>(v, k);
    for (int j = 0; j < first; j++) booth_reg_shift<8>(v, W);
    for (int i = first; i < nwin; i += step) {
        u32 neg;
        u32 d = booth_from_view(v[0], W, neg);
        for (int j = 0; j < step; j++) booth_reg_shift<8>(v, W);
        ge_niels e;
        ge_niels_identity(e);
        if (d != 0) {
            const u32* src = table + ((size_t)i * half + (d - 1)) * (size_t)stride;
            ld_words<8>(e.yp.v, src);
            ld_words<8>(e.ym.v, src + 8);
            ld_words<8>(e.t2d.v, src + 16);
        }
        ge_niels_cneg(e, neg);
        // first window: identity + entry needs one product; last window: T only if somebody reads it
        if (i == first) ge_from_niels(acc, e);
        else ge_madd_rt(acc, acc, e, want_t || (i + step < nwin));
    }
}
template <bool CLAMP>
ECB_DEV void ed25519_mul_base_body(size_t idx, size_t n, const u32* scalars, const u32* table, int W, int nwin, int stride,
                                   u32* planes, unsigned long long* status) {
    u32 k[9];
    ed25519_load_scalar<CLAMP>(k, idx, scalars, status);
    ge_p3 acc;
    ed25519_comb_partial(acc, k, table, W, nwin, stride, 0, 1, false);
    plane_st<8>(planes + 0 * 8 * n, n, idx, acc.X.v);
    plane_st<8>(planes + 1 * 8 * n, n, idx, acc.Y.v);
    plane_st<8>(planes + 2 * 8 * n, n, idx, acc.Z.v);
}

// X25519 against the base point (protocol/x25519.rs:49 x25519_base = x25519(k, 9)): the ladder on
// u = 9 equals the u-coordinate of clamp(k) * B on the birationally equivalent edwards25519
// (u = (1 + y) / (1 - y), curve25519.rs:1668 tests this map), so the fixed-base comb replaces 255
// ladder steps.  Writes Z + Y -> plane 0, Z - Y -> plane 2; FinEdMontU divides (0 when Z = Y, as the
// ladder's z2 = 0 case gives).
ECB_DEV void x25519_base_body(size_t idx, size_t n, const u32* scalars, const u32* table, int W, int nwin, int stride, u32* planes) {
    ed25519_mul_base_body<true>(idx, n, scalars, table, W, nwin, stride, planes, nullptr);
    fe25519 Y, Z, s, d;
    plane_ld<8>(Y.v, planes + 1 * 8 * n, n, idx);
    plane_ld<8>(Z.v, planes + 2 * 8 * n, n, idx);
    F::add(s, Z, Y);
    F::sub(d, Z, Y);
    plane_st<8>(planes + 0 * 8 * n, n, idx, s.v);
    plane_st<8>(planes + 2 * 8 * n, n, idx, d.v);
}

// comb-table builder, two steps:
//   window bases  B_i = 2^(W i) * B, one thread per window (W i doublings), stored in cached form
//                 (Y+X, Y-X, Z, 2dT: 32 words) so that no inversion is needed in between;
//   entries       (i, j) = j * B_i by double-and-add over the W bits of j only (the first version
//                 walked all W * nwin bits of j << (W i) for every entry: ~10x the work).
// Output: projective planes (X, Y, Z) over ntab entries, made affine by the batch inversion.
ECB_DEV void ed25519_window_base_body(int i, int W, u32* bases) {
    fe25519 bx, by;
    F::from_words(bx, ED25519_BX);
    F::from_words(by, ED25519_BY);
    ge_p3 acc;
    ge_from_affine(acc, bx, by);
    for (int s = 0; s < W * i; s++) ge_double<true>(acc, acc);
    ge_cached c;
    ge_to_cached(c, acc);
    u32* dst = bases + (size_t)i * 32;
    st_words<8>(dst, c.yp.v);
    st_words<8>(dst + 8, c.ym.v);
    st_words<8>(dst + 16, c.Z.v);
    st_words<8>(dst + 24, c.t2d.v);
}
ECB_DEV void ed25519_table_point_body(size_t e, size_t ntab, int W, int nwin, const u32* bases, u32* planes) {
    const u32 half = 1u << (W - 1);
    u32 i = (u32)(e / half), j = (u32)(e % half) + 1;
    (void)nwin;
    ge_cached c;
    const u32* src = bases + (size_t)i * 32;
    ld_words<8>(c.yp.v, src);
    ld_words<8>(c.ym.v, src + 8);
    ld_words<8>(c.Z.v, src + 16);
    ld_words<8>(c.t2d.v, src + 24);
    ge_p3 acc;
    ge_identity(acc);
    for (int bit = W - 1; bit >= 0; bit--) {  // j <= 2^(W-1)
        ge_double<true>(acc, acc);
        if ((j >> bit) & 1) ge_add_cached<true>(acc, acc, c);
    }
    plane_st<8>(planes + 0 * 8 * ntab, ntab, e, acc.X.v);
    plane_st<8>(planes + 1 * 8 * ntab, ntab, e, acc.Y.v);
    plane_st<8>(planes + 2 * 8 * ntab, ntab, e, acc.Z.v);
}

// =======================================================================================
// Ed25519 variable base: k*P, signed 4-bit Booth windows over a per-thread table of 8 cached
// multiples (replaces &Point * &Scalar, curve25519.rs:746-760 / :1274).
//   tbl : this thread's scratch, 8 entries x 32 words
// =======================================================================================
ECB_DEV void ed25519_mul_body(size_t idx, size_t n, const u32* scalars, const u32* points, u32* tbl, u32* planes,
                              unsigned long long* status) {
    u32 k[8], w[16];
    ld_words<8>(k, scalars + idx * 8);
    ld_words<16>(w, points + idx * 16);
    u32 ok = 1;
    if (!lt_words8(k, ED25519_L)) {
        report_bad(status, idx, ST_NONCANONICAL_SCALAR);
        ok = 0;
    }
    fe25519 x, y;
    F::from_words(x, w);
    F::from_words(y, w + 8);
    if (ok && !(F::is_canonical_words(w) && F::is_canonical_words(w + 8) && ge_on_curve(x, y))) {
        report_bad(status, idx, ST_BAD_POINT);
        ok = 0;
    }
    if (!ok) {
        // neutral inputs keep the arithmetic defined; the host discards the batch on error
        ECB_UNROLL
        for (int i = 0; i < 8; i++) k[i] = 0;
        F::set_zero(x);
        F::set_one(y);
    }
    ge_p3 P, acc;
    ge_from_affine(P, x, y);
    // table[j-1] = cached(j*P), j = 1..8
    {
        ge_cached c, c1;
        ge_to_cached(c1, P);
        acc = P;
        ECB_NOUNROLL
        for (int j = 1; j <= 8; j++) {  // the addition is complete, so 2P = P + P needs no doubling
            ge_to_cached(c, acc);
            u32* d = tbl + (j - 1) * 32;
            st_words<8>(d + 0, c.yp.v); st_words<8>(d + 8, c.ym.v); st_words<8>(d + 16, c.Z.v); st_words<8>(d + 24, c.t2d.v);
            if (j < 8) ge_add_cached<true>(acc, acc, c1);
        }
    }
    // k < 2^253: 64 Booth windows of width 4 cover 256 bits
    ge_identity(acc);
    ECB_NOUNROLL
    for (int i = 63; i >= 0; i--) {
        if (i != 63) {
            ECB_NOUNROLL
            for (int r = 0; r < 4; r++) ge_double_rt<false>(acc, acc, r == 3 ? 1u : 0u);
        }
        u32 neg;
        u32 d = booth_digit(k, 8, 4, i, neg);
        ge_cached c;
        ge_cached_identity(c);
        if (d != 0) {
            const u32* s = tbl + (d - 1) * 32;
            ld_words_rw<8>(c.yp.v, s); ld_words_rw<8>(c.ym.v, s + 8); ld_words_rw<8>(c.Z.v, s + 16); ld_words_rw<8>(c.t2d.v, s + 24);
        }
        ge_cached_cneg(c, neg);
        ge_add_cached<true>(acc, acc, c);
    }
    plane_st<8>(planes + 0 * 8 * n, n, idx, acc.X.v);
    plane_st<8>(planes + 1 * 8 * n, n, idx, acc.Y.v);
    plane_st<8>(planes + 2 * 8 * n, n, idx, acc.Z.v);
}

// =======================================================================================
// X25519: the reference's ladder, step for step (protocol/x25519.rs:36-45, curve25519.rs:474-513):
// clamp, mask bit 255 of u (no canonical check), 256 ladder steps MSB first, final cswap.
// Writes x2 -> plane 0, z2 -> plane 2; u' = x2 * z2^-1 (0 for z2 = 0) is the batch inversion.
// =======================================================================================
ECB_DEV void x25519_body(size_t idx, size_t n, const u32* scalars, const u32* us, u32* planes) {
    u32 k[8], uw[8];
    ld_words<8>(k, scalars + idx * 8);
    ld_words<8>(uw, us + idx * 8);
    k[0] &= 0xfffffff8u;
    k[7] = (k[7] & 0x7fffffffu) | 0x40000000u;
    uw[7] &= 0x7fffffffu;
    fe25519 x1, x2, z2, x3, z3;
    F::from_words(x1, uw);
    F::set_one(x2);
    F::set_zero(z2);
    F::copy(x3, x1);
    F::set_one(z3);
    u32 swap = 0;
    for (int wi = 7; wi >= 0; wi--) {
        u32 word = 0;
        ECB_UNROLL
        for (int t = 0; t < 8; t++)
            if (t == wi) word = k[t];
        for (int bi = 31; bi >= 0; bi--) {
            u32 bit = (word >> bi) & 1u;
            swap ^= bit;
            F::cswap(swap, x2, x3);
            F::cswap(swap, z2, z3);
            swap = bit;
            fe25519 a, aa, b, bb, e, c, d, da, cb, t;
            F::add(a, x2, z2);
            F::sqr(aa, a);
            F::sub(b, x2, z2);
            F::sqr(bb, b);
            F::sub(e, aa, bb);
            F::add(c, x3, z3);
            F::sub(d, x3, z3);
            F::mul(da, d, a);
            F::mul(cb, c, b);
            F::add(t, da, cb);
            F::sqr(x3, t);
            F::sub(t, da, cb);
            F::sqr(t, t);
            F::mul(z3, x1, t);
            F::mul(x2, aa, bb);
            F::mul_small(t, e, 121666u);
            F::add(t, bb, t);
            F::mul(z2, e, t);
        }
    }
    F::cswap(swap, x2, x3);
    F::cswap(swap, z2, z3);
    plane_st<8>(planes + 0 * 8 * n, n, idx, x2.v);
    plane_st<8>(planes + 2 * 8 * n, n, idx, z2.v);
}

// =======================================================================================
// Batch inversion (Montgomery's trick) + affine finishing.
//   thread t owns elements t, t+T, t+2T, ... ; prefix products go to the `pf` plane.
//   Z == 0 is replaced by 1 and reported to the finisher (Weierstrass infinity,
//   X25519 z2 = 0 -> output 0 as invert_or_zero does, curve25519.rs:191, :512).
// Cost per element: 3 M + the finisher's (2 M for x,y), plus one inversion per thread.
// =======================================================================================
// forward pass: running products of the thread's elements (t, t+T, t+2T, ...) into `pf`, two
// interleaved chains (elements t, t+2T, ... and t+T, t+3T, ...): the two running products are
// independent, so the dependent multiplication chains overlap; returns the two chain totals.
template <class FT>
ECB_DEV void batch_inv_forward(size_t t, size_t T, size_t n, const u32* planes, u32* pf,
                               typename FT::el& accA, typename FT::el& accB) {
    typedef typename FT::el fe;
    constexpr int N = FT::N;
    const u32* zp = planes + 2 * (size_t)N * n;
    fe zA, zB, one;
    FT::set_one(one);
    FT::set_one(accA);
    FT::set_one(accB);
    size_t cnt = t < n ? (n - t + T - 1) / T : 0;  // elements owned by this thread
    size_t pairs = cnt / 2;
    for (size_t j = 0; j < pairs; j++) {
        size_t ia = t + (2 * j) * T, ib = ia + T;
        if (ib + 2 * T < n) {                    // the loop is bound by DRAM latency without this
            plane_prefetch<N>(zp, n, ia + 2 * T);
            plane_prefetch<N>(zp, n, ib + 2 * T);
        }
        plane_ld<N>(zA.v, zp, n, ia);
        plane_ld<N>(zB.v, zp, n, ib);
        u32 zeroA = FT::is_zero(zA), zeroB = FT::is_zero(zB);
        FT::select(zA, zeroA, one, zA);
        FT::select(zB, zeroB, one, zB);
        plane_st<N>(pf, n, ia, accA.v);
        plane_st<N>(pf, n, ib, accB.v);
        FT::mul(accA, accA, zA);
        FT::mul(accB, accB, zB);
    }
    if (cnt & 1) {                               // odd leftover goes to chain A
        size_t ia = t + (cnt - 1) * T;
        plane_ld<N>(zA.v, zp, n, ia);
        u32 zeroA = FT::is_zero(zA);
        FT::select(zA, zeroA, one, zA);
        plane_st<N>(pf, n, ia, accA.v);
        FT::mul(accA, accA, zA);
    }
}
// backward pass: invA / invB are the inverses of the two chain totals
template <class FT, class FIN>
ECB_DEV void batch_inv_backward(size_t t, size_t T, size_t n, const u32* planes, const u32* pf, FIN fin,
                                typename FT::el invA, typename FT::el invB) {
    typedef typename FT::el fe;
    constexpr int N = FT::N;
    const u32* zp = planes + 2 * (size_t)N * n;
    fe zA, zB, one;
    FT::set_one(one);
    size_t cnt = t < n ? (n - t + T - 1) / T : 0;
    size_t pairs = cnt / 2;
    if (cnt & 1) {
        size_t ia = t + (cnt - 1) * T;
        plane_ld<N>(zA.v, zp, n, ia);
        u32 zeroA = FT::is_zero(zA);
        FT::select(zA, zeroA, one, zA);
        fe p, zinv;
        plane_ld<N>(p.v, pf, n, ia);
        FT::mul(zinv, invA, p);
        FT::mul(invA, invA, zA);
        fin(ia, zinv, zeroA);
    }
    for (size_t j = pairs; j-- > 0;) {
        size_t ia = t + (2 * j) * T, ib = ia + T;
        if (j > 0) {
            size_t pa = ia - 2 * T, pb = ib - 2 * T;
            plane_prefetch<N>(zp, n, pa);
            plane_prefetch<N>(zp, n, pb);
            plane_prefetch<N>(pf, n, pa);
            plane_prefetch<N>(pf, n, pb);
            fin.pre(pa);
            fin.pre(pb);
        }
        plane_ld<N>(zA.v, zp, n, ia);
        plane_ld<N>(zB.v, zp, n, ib);
        u32 zeroA = FT::is_zero(zA), zeroB = FT::is_zero(zB);
        FT::select(zA, zeroA, one, zA);
        FT::select(zB, zeroB, one, zB);
        fe pA, pB, zinvA, zinvB;
        plane_ld<N>(pA.v, pf, n, ia);
        plane_ld<N>(pB.v, pf, n, ib);
        FT::mul(zinvA, invA, pA);
        FT::mul(zinvB, invB, pB);
        FT::mul(invA, invA, zA);
        FT::mul(invB, invB, zB);
        fin(ia, zinvA, zeroA);
        fin(ib, zinvB, zeroB);
    }
}
// one inversion per thread (the host simulation and option inv_block = 0 use this form; the
// block-cooperative kernel in tu_common.cuh shares the two passes and inverts once per block)
// CT: the inverse in the middle is the fixed Fermat chain instead of the variable-time safegcd (secret data)
template <class FT, class FIN, bool CT = false>
ECB_DEV void batch_inv_body(size_t t, size_t T, size_t n, const u32* planes, u32* pf, FIN fin) {
    typedef typename FT::el fe;
    if (t >= n) return;
    fe accA, accB, inv, invA, invB;
    batch_inv_forward<FT>(t, T, n, planes, pf, accA, accB);
    FT::mul(inv, accA, accB);
    if (CT) FT::invert_fermat(inv, inv);
    else FT::invert(inv, inv);
    FT::mul(invA, inv, accB);
    FT::mul(invB, inv, accA);
    batch_inv_backward<FT, FIN>(t, T, n, planes, pf, fin, invA, invB);
}

// finishers ------------------------------------------------------------------------------
struct FinEdXY {  // out: x_le || y_le, canonical (Point::to_affine + to_bytes_le)
    const u32* planes; size_t n; u32* out;
    ECB_DEV void pre(size_t idx) const { plane_prefetch<8>(planes, n, idx); plane_prefetch<8>(planes + 8 * n, n, idx); }
    ECB_DEV void operator()(size_t idx, const fe25519& zinv, u32) const {
        fe25519 X, Y, x, y;
        plane_ld<8>(X.v, planes, n, idx);
        plane_ld<8>(Y.v, planes + 8 * n, n, idx);
        F::mul(x, X, zinv);
        F::mul(y, Y, zinv);
        F::freeze(x, x);
        F::freeze(y, y);
        st_words<8>(out + idx * 16, x.v);
        st_words<8>(out + idx * 16 + 8, y.v);
    }
};
struct FinEdCompressed {  // out: encode_point (protocol/ed25519.rs:27); element idx at out + idx * stride words
    const u32* planes; size_t n; u32* out; size_t stride = 8;
    ECB_DEV void pre(size_t idx) const { plane_prefetch<8>(planes, n, idx); plane_prefetch<8>(planes + 8 * n, n, idx); }
    ECB_DEV void operator()(size_t idx, const fe25519& zinv, u32) const {
        fe25519 X, Y, x, y;
        plane_ld<8>(X.v, planes, n, idx);
        plane_ld<8>(Y.v, planes + 8 * n, n, idx);
        F::mul(x, X, zinv);
        F::mul(y, Y, zinv);
        F::freeze(x, x);
        F::freeze(y, y);
        y.v[7] |= (x.v[0] & 1u) << 31;
        st_words<8>(out + idx * stride, y.v);
    }
};
struct FinEdMontU {  // out: u = (1 + y) / (1 - y) = (Z + Y) / (Z - Y), little-endian canonical; 0 for the identity
    const u32* planes; size_t n; u32* out;   // the "Z" plane holds Z - Y (see ed25519_to_montgomery_planes)
    ECB_DEV void pre(size_t idx) const { plane_prefetch<8>(planes, n, idx); }
    ECB_DEV void operator()(size_t idx, const fe25519& dinv, u32 zero) const {
        fe25519 N_, u;
        plane_ld<8>(N_.v, planes, n, idx);
        F::mul(u, N_, dinv);
        F::freeze(u, u);
        u32 m = zero ? 0u : 0xffffffffu;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) u.v[i] &= m;
        st_words<8>(out + idx * 8, u.v);
    }
};
struct FinEdNiels {  // out: comb-table entry (y+x, y-x, 2dxy), 24 words at out + idx * stride
    const u32* planes; size_t n; u32* out; size_t stride = 24;
    ECB_DEV void pre(size_t idx) const { plane_prefetch<8>(planes, n, idx); plane_prefetch<8>(planes + 8 * n, n, idx); }
    ECB_DEV void operator()(size_t idx, const fe25519& zinv, u32) const {
        fe25519 X, Y, x, y;
        plane_ld<8>(X.v, planes, n, idx);
        plane_ld<8>(Y.v, planes + 8 * n, n, idx);
        F::mul(x, X, zinv);
        F::mul(y, Y, zinv);
        ge_niels e;
        ge_niels_from_affine(e, x, y);
        F::freeze(e.yp, e.yp);
        F::freeze(e.ym, e.ym);
        F::freeze(e.t2d, e.t2d);
        st_words<8>(out + idx * stride, e.yp.v);
        st_words<8>(out + idx * stride + 8, e.ym.v);
        st_words<8>(out + idx * stride + 16, e.t2d.v);
    }
};
struct FinX25519 {  // out: u' = x2 / z2 little-endian canonical, 0 when z2 == 0
    const u32* planes; size_t n; u32* out;
    ECB_DEV void pre(size_t idx) const { plane_prefetch<8>(planes, n, idx); }
    ECB_DEV void operator()(size_t idx, const fe25519& zinv, u32 zero) const {
        fe25519 X, x;
        plane_ld<8>(X.v, planes, n, idx);
        F::mul(x, X, zinv);
        F::freeze(x, x);
        u32 m = zero ? 0u : 0xffffffffu;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) x.v[i] &= m;
        st_words<8>(out + idx * 8, x.v);
    }
};

// =======================================================================================
// Weierstrass variable base: k*P with signed Booth windows of C::WIN bits (4, or 5 on p256r1) over a per-thread
// table of 2^(WIN - 1) Jacobian multiples (replaces &Point * &Scalar, fiat/curve_macros.rs:321 ->
// projective.rs:871 scalar_mul_fixed_window_am3 / :842 _a0: same group element, different
// coordinates and window recoding — only the canonical affine result is observable).
//   scalars: n x SB bytes big-endian canonical (< group order)
//   points : n x 2FB bytes big-endian affine (x, y), on the curve; inf_in (optional) marks
//            identity inputs
//   tbl    : this thread's scratch, 2^(C::WIN - 1) entries x 5N words (X, Y, Z, Z^2, Z^3)
// =======================================================================================
// tbl[j-1] = cached(j * (x, y)), j = 1..2^(WIN - 1): half of them doublings, the rest additions of P.  Entries are written to
// the table as soon as they exist and re-read when needed, so a single point is live at a time.
template <class C>
ECB_DEV void wei_store_pt_cached(u32* d, const typename WeiJ<C>::pt& p) {
    typedef typename C::F FT;
    constexpr int N = FT::N;
    typename FT::el zz, zzz;
    FT::sqr_ni(zz, p.Z);
    FT::mul_ni(zzz, zz, p.Z);
    st_words<N>(d, p.X.v); st_words<N>(d + N, p.Y.v); st_words<N>(d + 2 * N, p.Z.v);
    st_words<N>(d + 3 * N, zz.v); st_words<N>(d + 4 * N, zzz.v);
}
// HALF = 2^(WIN - 1) entries, WIN = C::WIN the width of the signed windows (weier.cuh)
template <class C>
ECB_DEV void wei_build_table(u32* tbl, const typename C::F::el& x, const typename C::F::el& y) {
    typedef WeiJ<C> J;
    typedef typename C::F FT;
    constexpr int N = FT::N;
    constexpr int ES = 5 * N;
    auto ld = [](u32* dst, const u32* src) { ld_words_rw<N>(dst, src); };
    typename J::pt cur;
    FT::copy(cur.X, x); FT::copy(cur.Y, y); FT::set_one(cur.Z);
    st_words<N>(tbl, cur.X.v); st_words<N>(tbl + N, cur.Y.v);
    st_words<N>(tbl + 2 * N, cur.Z.v); st_words<N>(tbl + 3 * N, cur.Z.v); st_words<N>(tbl + 4 * N, cur.Z.v);
    J::dbl(cur, cur);                                   // 2P
    wei_store_pt_cached<C>(tbl + 1 * ES, cur);
    ECB_NOUNROLL
    for (int j = 2; j <= WeiJ<C>::TBL - 2; j += 2) {    // (j+1)P = jP + P ; (j+2)P = 2 * ((j+2)/2)P
        J::add_mem(cur, cur, tbl, 0u, ld);
        wei_store_pt_cached<C>(tbl + j * ES, cur);
        const u32* h = tbl + (j / 2) * ES;              // ((j+2)/2) P sits in slot (j+2)/2 - 1 = j/2
        ld(cur.X.v, h); ld(cur.Y.v, h + N); ld(cur.Z.v, h + 2 * N);
        J::dbl(cur, cur);
        wei_store_pt_cached<C>(tbl + (j + 1) * ES, cur);
    }
}

// acc = k * P over the table of multiples at tbl: signed windows of C::WIN bits, most significant first
template <class C, bool LZ>
ECB_DEV void wei_window_loop(typename WeiJ<C>::pt& acc, const u32* k, const u32* tbl, typename C::F::lazy& z) {
    typedef WeiJ<C> J;
    constexpr int N = C::F::N;
    constexpr int NS = C::SB / 4;
    constexpr int ES = 5 * N;
    constexpr int WIN = C::WIN;
    constexpr int NWIN = (C::SBITS + 1 + WIN - 1) / WIN;
    J::set_inf(acc);
    ECB_NOUNROLL
    for (int i = NWIN - 1; i >= 0; i--) {
        if (i != NWIN - 1) {
            ECB_NOUNROLL
            for (int r = 0; r < WIN; r++) J::template dbl_z<LZ>(acc, acc, z);
        }
        u32 neg;
        u32 d = booth_digit(k, NS + 1, WIN, i, neg);
        if (d != 0) J::add_mem(acc, acc, tbl + (d - 1) * ES, neg, [](u32* dst, const u32* src) { ld_words_rw<N>(dst, src); });
    }
}
template <class C>
ECB_DEVNI typename WeiJ<C>::pt wei_window_loop_checked(const u32* k, const u32* tbl) {
    typename WeiJ<C>::pt acc;
    typename C::F::lazy z;
    wei_window_loop<C, false>(acc, k, tbl, z);
    return acc;
}
template <class C>
ECB_DEV void wei_mul_body(size_t idx, size_t n, const u32* scalars, const u32* points, const unsigned char* inf_in,
                          u32* tbl, u32* planes, unsigned long long* status) {
    typedef WeiJ<C> J;
    typedef Wei<C> W;
    typedef typename C::F FT;
    typedef typename C::FN FNT;
    typedef typename FT::el fe;
    constexpr int N = FT::N;
    constexpr int NS = C::SB / 4;
    constexpr int ES = 5 * N;
    u32 k[NS + 1];
    ld_words_be<NS>(k, scalars + idx * NS);
    k[NS] = 0;
    u32 ok = 1;
    if (!FNT::is_canonical_words(k)) {
        report_bad(status, idx, ST_NONCANONICAL_SCALAR);
        ok = 0;
    }
    u32 is_inf = inf_in ? (inf_in[idx] ? 1u : 0u) : 0u;
    fe px, py;
    {
        u32 xw[N], yw[N];
        ld_words_be<N>(xw, points + idx * 2 * N);
        ld_words_be<N>(yw, points + idx * 2 * N + N);
        FT::to_mont(px, xw);
        FT::to_mont(py, yw);
        if (ok && !is_inf && !(FT::is_canonical_words(xw) && FT::is_canonical_words(yw) && W::on_curve(px, py))) {
            report_bad(status, idx, ST_BAD_POINT);
            ok = 0;
        }
    }
    typename J::pt acc;
    J::set_inf(acc);
    if (ok && !is_inf) {
        wei_build_table<C>(tbl, px, py);
        // the doublings run on the lazy field forms (mont.cuh): a fold that carries again is only counted, and the whole
        // window loop is then redone with the checked forms — about one element in 2^20 on the loose fields
        typename FT::lazy z;
        wei_window_loop<C, FT::LOOSE>(acc, k, tbl, z);
        if (z.any()) acc = wei_window_loop_checked<C>(k, tbl);
    }
    plane_st<N>(planes + 0 * (size_t)N * n, n, idx, acc.X.v);
    plane_st<N>(planes + 1 * (size_t)N * n, n, idx, acc.Y.v);
    plane_st<N>(planes + 2 * (size_t)N * n, n, idx, acc.Z.v);
}

// =======================================================================================
// Weierstrass fixed base: k*G from a signed-digit comb (replaces Point::mul_base,
// fiat/curve_macros.rs:55 -> projective.rs:965 mul_base_table_am3 / :945 _a0 and the tables of
// src/params/comb/{p256r1,p384r1,bls12_381}.rs — regenerated on the device, wider windows):
//   table[(i * half + (j-1)) * 2N ..] = (x, y) of j * 2^(W i) * G in the Montgomery domain,
//   j = 1..half = 2^(W-1), i = 0..nwin-1, nwin = ceil((SBITS + 1) / W).
// One mixed Jacobian addition (7M + 4S) per window, no doublings.
// =======================================================================================
template <class C>
ECB_DEV void wei_comb_accumulate(typename WeiJ<C>::pt& acc, const u32* k, int nwords, const u32* table, int W, int nwin) {
    typedef WeiJ<C> J;
    typedef typename C::F FT;
    constexpr int N = FT::N;
    const u32 half = 1u << (W - 1);
    constexpr int NV = C::SB / 4 + 1;   // callers pass k[NS + 1] with a zero top word
    (void)nwords;
    u32 v[NV];
    booth_reg_init<NV>(v, k);
    ECB_NOUNROLL
    for (int i = 0; i < nwin; i++) {
        u32 neg;
        u32 d = booth_from_view(v[0], W, neg);
        booth_reg_shift<NV>(v, W);
        if (d != 0) {
            typename J::cached e;
            const u32* src = table + ((size_t)i * half + (d - 1)) * 2 * N;
            ld_words<N>(e.X.v, src);
            ld_words<N>(e.Y.v, src + N);
            J::cached_cneg(e, neg);
            J::template add<true>(acc, acc, e);
        }
    }
}
template <class C>
ECB_DEV void wei_mul_base_body(size_t idx, size_t n, const u32* scalars, const u32* table, int W, int nwin, u32* planes,
                               unsigned long long* status) {
    typedef WeiJ<C> J;
    typedef typename C::FN FNT;
    constexpr int N = C::F::N;
    constexpr int NS = C::SB / 4;
    u32 k[NS + 1];
    ld_words_be<NS>(k, scalars + idx * NS);
    k[NS] = 0;
    if (!FNT::is_canonical_words(k)) {
        report_bad(status, idx, ST_NONCANONICAL_SCALAR);
        ECB_UNROLL
        for (int i = 0; i < NS; i++) k[i] = 0;
    }
    typename J::pt acc;
    J::set_inf(acc);
    wei_comb_accumulate<C>(acc, k, NS + 1, table, W, nwin);
    plane_st<N>(planes + 0 * (size_t)N * n, n, idx, acc.X.v);
    plane_st<N>(planes + 1 * (size_t)N * n, n, idx, acc.Y.v);
    plane_st<N>(planes + 2 * (size_t)N * n, n, idx, acc.Z.v);
}
// comb-table builder, as for Ed25519: window bases G_i = 2^(W i) * G (Jacobian with cached Z^2, Z^3:
// 5N words each), then entry (i, j) = j * G_i over the W bits of j.
template <class C>
ECB_DEV void wei_window_base_body(int i, int W, u32* bases) {
    typedef WeiJ<C> J;
    typedef typename C::F FT;
    constexpr int N = FT::N;
    typename J::pt acc;
    ECB_UNROLL
    for (int t = 0; t < N; t++) { acc.X.v[t] = C::gx(t); acc.Y.v[t] = C::gy(t); }
    FT::set_one(acc.Z);
    ECB_NOUNROLL
    for (int s = 0; s < W * i; s++) J::dbl(acc, acc);
    typename J::cached c;
    J::to_cached(c, acc);
    u32* dst = bases + (size_t)i * 5 * N;
    st_words<N>(dst, c.X.v);
    st_words<N>(dst + N, c.Y.v);
    st_words<N>(dst + 2 * N, c.Z.v);
    st_words<N>(dst + 3 * N, c.ZZ.v);
    st_words<N>(dst + 4 * N, c.ZZZ.v);
}
template <class C>
ECB_DEV void wei_table_point_body(size_t e, size_t ntab, int W, int nwin, const u32* bases, u32* planes) {
    typedef WeiJ<C> J;
    typedef typename C::F FT;
    constexpr int N = FT::N;
    const u32 half = 1u << (W - 1);
    u32 i = (u32)(e / half), j = (u32)(e % half) + 1;
    (void)nwin;
    typename J::cached g;
    const u32* src = bases + (size_t)i * 5 * N;
    ld_words<N>(g.X.v, src);
    ld_words<N>(g.Y.v, src + N);
    ld_words<N>(g.Z.v, src + 2 * N);
    ld_words<N>(g.ZZ.v, src + 3 * N);
    ld_words<N>(g.ZZZ.v, src + 4 * N);
    typename J::pt acc;
    J::set_inf(acc);
    ECB_NOUNROLL
    for (int bit = W - 1; bit >= 0; bit--) {  // j <= 2^(W-1)
        J::dbl(acc, acc);
        if ((j >> bit) & 1) J::template add<false>(acc, acc, g);
    }
    plane_st<N>(planes + 0 * (size_t)N * ntab, ntab, e, acc.X.v);
    plane_st<N>(planes + 1 * (size_t)N * ntab, ntab, e, acc.Y.v);
    plane_st<N>(planes + 2 * (size_t)N * ntab, ntab, e, acc.Z.v);
}
template <class C>
struct FinWeiTable {  // out: comb-table entry (x, y), Montgomery domain, canonical, 2N words
    typedef typename C::F FT;
    const u32* planes; size_t n; u32* out;
    ECB_DEV void pre(size_t idx) const { plane_prefetch<FT::N>(planes, n, idx); plane_prefetch<FT::N>(planes + (size_t)FT::N * n, n, idx); }
    ECB_DEV void operator()(size_t idx, const typename FT::el& zinv, u32) const {
        constexpr int N = FT::N;
        typename FT::el X, Y, x, y, zi2;
        plane_ld<N>(X.v, planes, n, idx);
        plane_ld<N>(Y.v, planes + (size_t)N * n, n, idx);
        FT::sqr(zi2, zinv);
        FT::mul(x, X, zi2);
        FT::mul(zi2, zi2, zinv);
        FT::mul(y, Y, zi2);
        FT::canon(x, x);
        FT::canon(y, y);
        st_words<N>(out + idx * 2 * N, x.v);
        st_words<N>(out + idx * 2 * N + N, y.v);
    }
};

template <class C>
struct FinWeiXY {  // out: x_be || y_be (canonical, out of the Montgomery domain) + infinity flag
    typedef typename C::F FT;
    const u32* planes; size_t n; u32* out; unsigned char* inf_out;
    ECB_DEV void pre(size_t idx) const { plane_prefetch<FT::N>(planes, n, idx); plane_prefetch<FT::N>(planes + (size_t)FT::N * n, n, idx); }
    ECB_DEV void operator()(size_t idx, const typename FT::el& zinv, u32 zero) const {
        constexpr int N = FT::N;
        typename FT::el X, Y, x, y;
        plane_ld<N>(X.v, planes, n, idx);
        plane_ld<N>(Y.v, planes + (size_t)N * n, n, idx);
        typename FT::el zi2;
        FT::sqr(zi2, zinv);          // Jacobian: x = X / Z^2, y = Y / Z^3
        FT::mul(x, X, zi2);
        FT::mul(zi2, zi2, zinv);
        FT::mul(y, Y, zi2);
        u32 xw[N], yw[N];
        FT::from_mont(xw, x);
        FT::from_mont(yw, y);
        u32 m = zero ? 0u : 0xffffffffu;
        ECB_UNROLL
        for (int i = 0; i < N; i++) { xw[i] &= m; yw[i] &= m; }
        st_words_be<N>(out + idx * 2 * N, xw);
        st_words_be<N>(out + idx * 2 * N + N, yw);
        if (inf_out) inf_out[idx] = (unsigned char)zero;
    }
};

}  // namespace ecb
