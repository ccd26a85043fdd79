// edwards.cuh — edwards25519 group arithmetic (extended coordinates, a = -1) and the
// curve25519 x-only ladder, device side.
//
// Group law = the reference's Point (src/curve/curve25519.rs:592-757): double_parts :604,
// add :695 (HWCD complete addition with 2d), add_cached :715.  The batch path is free to
// schedule the group operations differently (signed windows, precomputed affine "niels"
// entries with Z = 1) because only canonical affine bytes are observable.
#pragma once
#include "fe25519.cuh"

namespace ecb {

typedef F25519 F;

struct ge_p3 {
    fe25519 X, Y, Z, T;
};
// affine precomputed point: (y+x, y-x, 2d*x*y)
struct ge_niels {
    fe25519 yp, ym, t2d;
};
// projective precomputed point: (Y+X, Y-X, Z, 2d*T)  (CachedPoint, curve25519.rs:994)
struct ge_cached {
    fe25519 yp, ym, Z, t2d;
};

ECB_DEV void ge_identity(ge_p3& r) {
    F::set_zero(r.X);
    F::set_one(r.Y);
    F::set_one(r.Z);
    F::set_zero(r.T);
}
ECB_DEV void ge_from_affine(ge_p3& r, const fe25519& x, const fe25519& y) {
    F::copy(r.X, x);
    F::copy(r.Y, y);
    F::set_one(r.Z);
    F::mul(r.T, x, y);
}
// a*x^2 + y^2 == 1 + d*x^2*y^2 (Point::from_coordinate, curve25519.rs:649)
ECB_DEV u32 ge_on_curve(const fe25519& x, const fe25519& y) {
    fe25519 xx, yy, l, r, d, one;
    F::sqr(xx, x);
    F::sqr(yy, y);
    F::sub(l, yy, xx);
    F::from_words(d, ED25519_D);
    F::mul(r, xx, yy);
    F::mul(r, r, d);
    F::set_one(one);
    F::add(r, r, one);
    return F::eq(l, r);
}

// r = 2p. 4S + 3M (+1M when WITH_T)
template <bool WITH_T>
ECB_DEV void ge_double(ge_p3& r, const ge_p3& p) {
    fe25519 A, B, C, E, G, Fv, H, t;
    F::sqr(A, p.X);
    F::sqr(B, p.Y);
    F::sqr(C, p.Z);
    F::dbl(C, C);
    F::add(t, p.X, p.Y);
    F::sqr(E, t);
    F::add(H, A, B);      // A + B
    F::sub(E, E, H);      // E = (X+Y)^2 - A - B
    F::sub(G, B, A);      // G = D + B = B - A
    F::sub(Fv, G, C);     // F = G - C
    F::neg(H, H);         // H = D - B = -(A + B)
    F::mul(r.X, E, Fv);
    F::mul(r.Y, G, H);
    F::mul(r.Z, Fv, G);
    if (WITH_T) F::mul(r.T, E, H);
}

// Same doubling with the T product behind a (warp-uniform) run-time flag: the window loops of the
// variable-base kernels keep ONE copy of the doubling in a 4-iteration loop instead of two inlined
// instances, which keeps the loop body inside the instruction cache.
// NI: field products through the out-of-line by-value copies (smaller loop body; pays off in the
// verification kernel, which is the largest, and costs ~9 % in the plain variable-base kernel).
template <bool NI>
ECB_DEV void ge_double_rt(ge_p3& r, const ge_p3& p, u32 with_t) {
    fe25519 A, B, C, E, G, Fv, H, t;
    if (NI) F::sqr_ni(A, p.X); else F::sqr(A, p.X);
    if (NI) F::sqr_ni(B, p.Y); else F::sqr(B, p.Y);
    if (NI) F::sqr_ni(C, p.Z); else F::sqr(C, p.Z);
    F::dbl(C, C);
    F::add(t, p.X, p.Y);
    if (NI) F::sqr_ni(E, t); else F::sqr(E, t);
    F::add(H, A, B);
    F::sub(E, E, H);
    F::sub(G, B, A);
    F::sub(Fv, G, C);
    F::neg(H, H);
    if (NI) F::mul_ni(r.X, E, Fv); else F::mul(r.X, E, Fv);
    if (NI) F::mul_ni(r.Y, G, H); else F::mul(r.Y, G, H);
    if (NI) F::mul_ni(r.Z, Fv, G); else F::mul(r.Z, Fv, G);
    if (with_t) { if (NI) F::mul_ni(r.T, E, H); else F::mul(r.T, E, H); }
}

// r = p + q, q affine precomputed. 7M (6M when !WITH_T)
template <bool WITH_T>
ECB_DEV void ge_madd(ge_p3& r, const ge_p3& p, const ge_niels& q) {
    fe25519 A, B, C, D, E, Fv, G, H;
    F::sub(A, p.Y, p.X);
    F::mul(A, A, q.ym);
    F::add(B, p.Y, p.X);
    F::mul(B, B, q.yp);
    F::mul(C, p.T, q.t2d);
    F::dbl(D, p.Z);
    F::sub(E, B, A);
    F::sub(Fv, D, C);
    F::add(G, D, C);
    F::add(H, B, A);
    F::mul(r.X, E, Fv);
    F::mul(r.Y, G, H);
    F::mul(r.Z, Fv, G);
    if (WITH_T) F::mul(r.T, E, H);
}

// the same with the T output decided at run time (one code body for every window of the comb loop)
ECB_DEV void ge_madd_rt(ge_p3& r, const ge_p3& p, const ge_niels& q, bool with_t) {
    fe25519 A, B, C, D, E, Fv, G, H;
    F::sub(A, p.Y, p.X);
    F::mul(A, A, q.ym);
    F::add(B, p.Y, p.X);
    F::mul(B, B, q.yp);
    F::mul(C, p.T, q.t2d);
    F::dbl(D, p.Z);
    F::sub(E, B, A);
    F::sub(Fv, D, C);
    F::add(G, D, C);
    F::add(H, B, A);
    F::mul(r.X, E, Fv);
    F::mul(r.Y, G, H);
    F::mul(r.Z, Fv, G);
    if (with_t) F::mul(r.T, E, H);
}
// r = identity + q: the addition formulas with (X, Y, Z, T) = (0, 1, 1, 0) leave
// E = yp - ym, H = yp + ym, F = G = 2, so r = (2E : 2H : 4 : E H) — one product instead of seven.
// The niels identity (1, 1, 0) gives (0 : 4 : 4 : 0), the identity again.
ECB_DEV void ge_from_niels(ge_p3& r, const ge_niels& q) {
    fe25519 E, H, two;
    F::sub(E, q.yp, q.ym);
    F::add(H, q.yp, q.ym);
    F::dbl(r.X, E);
    F::dbl(r.Y, H);
    F::set_one(two);
    F::dbl(two, two);
    F::dbl(r.Z, two);
    F::mul(r.T, E, H);
}

// r = p + q, q projective cached. 8M
template <bool WITH_T, bool NI = false>
ECB_DEV void ge_add_cached(ge_p3& r, const ge_p3& p, const ge_cached& q) {
    fe25519 A, B, C, D, E, Fv, G, H;
    F::sub(A, p.Y, p.X);
    if (NI) F::mul_ni(A, A, q.ym); else F::mul(A, A, q.ym);
    F::add(B, p.Y, p.X);
    if (NI) F::mul_ni(B, B, q.yp); else F::mul(B, B, q.yp);
    if (NI) F::mul_ni(C, p.T, q.t2d); else F::mul(C, p.T, q.t2d);
    if (NI) F::mul_ni(D, p.Z, q.Z); else F::mul(D, p.Z, q.Z);
    F::dbl(D, D);
    F::sub(E, B, A);
    F::sub(Fv, D, C);
    F::add(G, D, C);
    F::add(H, B, A);
    if (NI) F::mul_ni(r.X, E, Fv); else F::mul(r.X, E, Fv);
    if (NI) F::mul_ni(r.Y, G, H); else F::mul(r.Y, G, H);
    if (NI) F::mul_ni(r.Z, Fv, G); else F::mul(r.Z, Fv, G);
    if (WITH_T) { if (NI) F::mul_ni(r.T, E, H); else F::mul(r.T, E, H); }
}

// r = p + q, both extended (Point::add, curve25519.rs:695: 9M, 8M when !WITH_T).  Used by the
// lane-split comb: the partial sums of the lanes of one scalar are added pairwise.
template <bool WITH_T>
ECB_DEV void ge_add_p3(ge_p3& r, const ge_p3& p, const ge_p3& q) {
    fe25519 A, B, C, D, E, Fv, G, H, d2;
    F::from_words(d2, ED25519_D2);
    F::sub(A, p.Y, p.X);
    F::sub(B, q.Y, q.X);
    F::mul(A, A, B);
    F::add(B, p.Y, p.X);
    F::add(C, q.Y, q.X);
    F::mul(B, B, C);
    F::mul(C, p.T, q.T);
    F::mul(C, C, d2);
    F::mul(D, p.Z, q.Z);
    F::dbl(D, D);
    F::sub(E, B, A);
    F::sub(Fv, D, C);
    F::add(G, D, C);
    F::add(H, B, A);
    F::mul(r.X, E, Fv);
    F::mul(r.Y, G, H);
    F::mul(r.Z, Fv, G);
    if (WITH_T) F::mul(r.T, E, H);
}

ECB_DEV void ge_to_cached(ge_cached& r, const ge_p3& p) {
    fe25519 d2;
    F::from_words(d2, ED25519_D2);
    F::add(r.yp, p.Y, p.X);
    F::sub(r.ym, p.Y, p.X);
    F::copy(r.Z, p.Z);
    F::mul(r.t2d, p.T, d2);
}
ECB_DEV void ge_cached_identity(ge_cached& r) {
    F::set_one(r.yp);
    F::set_one(r.ym);
    F::set_one(r.Z);
    F::set_zero(r.t2d);
}
// conditional negation: (yp, ym, Z, t2d) -> (ym, yp, Z, -t2d)
ECB_DEV void ge_cached_cneg(ge_cached& r, u32 neg) {
    F::cswap(neg, r.yp, r.ym);
    fe25519 nt;
    F::neg(nt, r.t2d);
    F::select(r.t2d, neg, nt, r.t2d);
}
ECB_DEV void ge_niels_identity(ge_niels& r) {
    F::set_one(r.yp);
    F::set_one(r.ym);
    F::set_zero(r.t2d);
}
ECB_DEV void ge_niels_cneg(ge_niels& r, u32 neg) {
    F::cswap(neg, r.yp, r.ym);
    fe25519 nt;
    F::neg(nt, r.t2d);
    F::select(r.t2d, neg, nt, r.t2d);
}
// affine (x, y) -> niels
ECB_DEV void ge_niels_from_affine(ge_niels& r, const fe25519& x, const fe25519& y) {
    fe25519 d2, t;
    F::from_words(d2, ED25519_D2);
    F::add(r.yp, y, x);
    F::sub(r.ym, y, x);
    F::mul(t, x, y);
    F::mul(r.t2d, t, d2);
}

// Booth signed digit of width W for window i of the little-endian scalar words k[0..nw):
// d = -2^(W-1) b_{Wi+W-1} + sum_{j<W-1} 2^j b_{Wi+j} + b_{Wi-1}.  Returns |d| in [0, 2^(W-1)]
// and the sign.  sum_i d_i 2^(Wi) = k for ceil((bits+1)/W) windows.
ECB_DEV u32 booth_digit(const u32* k, int nwords, int W, int i, u32& neg) {
    int pos = W * i - 1;  // lowest bit of the (W+1)-bit view; -1 for i = 0
    u32 view;
    if (pos < 0) {
        view = k[0] << 1;
    } else {
        int wd = pos >> 5, sh = pos & 31;
        u32 lo = wd < nwords ? k[wd] : 0u;
        u32 hi = (wd + 1) < nwords ? k[wd + 1] : 0u;
        view = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
    }
    view &= (2u << W) - 1u;  // W+1 bits
    u32 s = view >> W;       // top bit = sign
    u32 d = (view + 1u) >> 1;
    // if sign: d = 2^W - d  (in units after the +1>>1 trick)
    u32 dn = (1u << W) - d;
    neg = s;
    u32 r = s ? dn : d;
    if (r == 0) neg = 0;
    return r;
}


// The comb kernels walk their windows in order, so instead of indexing k[] by a run-time word number
// (which sends the scalar to local memory) they keep v = 2k in a shift register: the (W+1)-bit Booth
// view of the current window is the low bits of v[0], and moving to the next window is a static
// multi-word funnel shift by W.  NV words hold 2k (the callers' scalars leave the top bit of the last
// word clear, or pass one word more).
ECB_DEV u32 funnel_r(u32 lo, u32 hi, int s) {   // low word of (hi:lo) >> s, 0 < s < 32
#ifdef ECB_HOSTSIM
    return (lo >> s) | (hi << (32 - s));
#else
    return __funnelshift_r(lo, hi, s);
#endif
}
template <int NV>
ECB_DEV void booth_reg_init(u32* v, const u32* k) {   // v = 2k
    v[0] = k[0] << 1;
    ECB_UNROLL
    for (int j = 1; j < NV; j++) v[j] = funnel_r(k[j - 1], k[j], 31);
}
template <int NV>
ECB_DEV void booth_reg_shift(u32* v, int W) {          // v >>= W, 0 < W < 32
    ECB_UNROLL
    for (int j = 0; j < NV - 1; j++) v[j] = funnel_r(v[j], v[j + 1], W);
    v[NV - 1] >>= W;
}
ECB_DEV u32 booth_from_view(u32 view, int W, u32& neg) {
    view &= (2u << W) - 1u;
    u32 s = view >> W;
    u32 d = (view + 1u) >> 1;
    u32 dn = (1u << W) - d;
    neg = s;
    u32 r = s ? dn : d;
    if (r == 0) neg = 0;
    return r;
}

}  // namespace ecb
