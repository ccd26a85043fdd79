// msm.cuh — multi-scalar multiplication sum_i k_i * P_i on a short-Weierstrass curve by the bucket method
// (Pippenger), SURVEY §8 f.4.  The reference lists it as a wish (TODO.md:48, :99-100); its per-element building
// blocks are Point::mul (projective.rs:842/871) and Point + Point (:268/:340), whose results this must reproduce:
// the output is the canonical affine encoding of the group element, as everywhere else on this path.
//
// With signed c-bit windows, k = sum_w d_w 2^(c w), |d_w| <= 2^(c-1):
//     sum_i k_i P_i = sum_w 2^(c w) * W_w,      W_w = sum_j j * B_(w,j),      B_(w,j) = sum of +-P_i with |d_(w,i)| = j.
// One scalar multiplication per point (256 doublings + 65 additions) becomes nwin = ceil((bits + 1) / c) mixed
// additions per point plus work that depends only on the number of buckets.  Kernels (one thread per ...):
//   msm_prepare_body   point        validate, convert to the Montgomery domain once, recode the scalar, count digits
//   (scan)             —            exclusive prefix of the counts = where each bucket's members start
//   msm_scatter_body   point        counting sort: write the point's index into each of its buckets
//   msm_bucket_body    bucket       B = sum of its members (mixed Jacobian additions, exceptional cases handled: the
//                                   same point may occur twice, P and -P may meet), then j * B by double-and-add
//   msm_reduce_body    bucket pair  tree sum of the j * B of one window (log2 levels)
//   msm_window_body    —            Horner over the windows: c doublings + one addition each
// Everything is integer work on the multiplier pipe, as the rest of the library; the gathers (96-byte points by
// index) stay in the L2 for a 2^20 batch.
#pragma once
#include "kernels.cuh"

namespace ecb {

template <class C>
struct Msm {
    typedef WeiJ<C> J;
    typedef typename C::F FT;
    typedef typename FT::el fe;
    static constexpr int N = FT::N;
    static constexpr int NS = C::SB / 4;
    static constexpr int PW = 3 * N;      // words of a stored Jacobian point (X, Y, Z)

    ECB_DEV static void st_pt(u32* dst, const typename J::pt& p) {
        st_words<N>(dst, p.X.v);
        st_words<N>(dst + N, p.Y.v);
        st_words<N>(dst + 2 * N, p.Z.v);
    }
    ECB_DEV static void ld_pt(typename J::pt& p, const u32* src) {
        ld_words_rw<N>(p.X.v, src);
        ld_words_rw<N>(p.Y.v, src + N);
        ld_words_rw<N>(p.Z.v, src + 2 * N);
    }
    // r = p + q for two Jacobian points (either may be infinity)
    ECB_DEV static void add_pts(typename J::pt& r, const typename J::pt& p, const typename J::pt& q) {
        if (J::is_inf(q)) {
            FT::copy(r.X, p.X); FT::copy(r.Y, p.Y); FT::copy(r.Z, p.Z);
            return;
        }
        typename J::cached cq;
        J::to_cached(cq, q);
        J::template add<false>(r, p, cq);
    }
};

// scalars n x SB bytes BE, points n x 2FB bytes BE.  pm: n x 2N words, the points in the Montgomery domain;
// dig: nwin x n signed digits (|d| in the low 31 bits, sign in bit 31); hist: nwin x NB counts, bucket j at slot j - 1.
template <class C>
ECB_DEV void msm_prepare_body(size_t i, size_t n, const u32* scalars, const u32* points, int c, int nwin, u32 NB, u32* pm, u32* dig,
                              u32* hist, unsigned long long* status) {
    typedef typename C::F FT;
    typedef typename C::FN FNT;
    typedef typename FT::el fe;
    constexpr int N = FT::N, NS = C::SB / 4, NV = NS + 1;
    u32 k[NV];
    ld_words_be<NS>(k, scalars + i * NS);
    k[NS] = 0;
    u32 ok = 1;
    if (!FNT::is_canonical_words(k)) {
        report_bad(status, i, ST_NONCANONICAL_SCALAR);
        ok = 0;
    }
    u32 xw[N], yw[N];
    ld_words_be<N>(xw, points + i * 2 * N);
    ld_words_be<N>(yw, points + i * 2 * N + N);
    fe x, y;
    FT::to_mont(x, xw);
    FT::to_mont(y, yw);
    if (ok && !(FT::is_canonical_words(xw) && FT::is_canonical_words(yw) && Wei<C>::on_curve(x, y))) {
        report_bad(status, i, ST_BAD_POINT);
        ok = 0;
    }
    st_words<N>(pm + i * 2 * N, x.v);
    st_words<N>(pm + i * 2 * N + N, y.v);
    const u32 mask = (1u << c) - 1u, half = 1u << (c - 1);
    u32 carry = 0;
    for (int w = 0; w < nwin; w++) {
        u32 raw = (k[0] & mask) + carry;
        booth_reg_shift<NV>(k, c);
        u32 neg = raw > half ? 1u : 0u;
        u32 d = neg ? (1u << c) - raw : raw;
        carry = neg;
        if (!ok) d = 0;
        dig[(size_t)w * n + i] = d | (d ? neg << 31 : 0u);
        if (d) atomicAdd(&hist[(size_t)w * NB + (d - 1)], 1u);
    }
}

// counting sort: the i-th point joins bucket (w, |d|); cursor starts at zero; idx entry = i | sign << 31
ECB_DEV void msm_scatter_body(size_t i, size_t n, int nwin, u32 NB, const u32* dig, const u32* offs, u32* cursor, u32* idx) {
    for (int w = 0; w < nwin; w++) {
        u32 e = dig[(size_t)w * n + i];
        u32 d = e & 0x7fffffffu;
        if (!d) continue;
        size_t b = (size_t)w * NB + (d - 1);
        u32 pos = atomicAdd(&cursor[b], 1u);
        idx[(size_t)offs[b] + pos] = (u32)i | (e & 0x80000000u);
    }
}

// bucket b = (w, j): B = sum of its members, then j * B -> bsum[b] (Jacobian, PW words)
template <class C>
ECB_DEV void msm_bucket_body(size_t b, u32 NB, const u32* offs, const u32* hist, const u32* idx, const u32* pm, u32* bsum) {
    typedef Msm<C> M;
    typedef typename M::J J;
    typedef typename C::F FT;
    constexpr int N = FT::N;
    const u32 j = (u32)(b % NB) + 1u;
    const u32 cnt = hist[b];
    const size_t start = offs[b];
    typename J::pt acc;
    J::set_inf(acc);
    ECB_NOUNROLL
    for (u32 t = 0; t < cnt; t++) {
        const u32 e = idx[start + t];
        const u32* src = pm + (size_t)(e & 0x7fffffffu) * 2 * N;
        typename J::cached q;
        ld_words<N>(q.X.v, src);
        ld_words<N>(q.Y.v, src + N);
        J::cached_cneg(q, e >> 31);
        J::template add<true>(acc, acc, q);
    }
    typename J::pt r;
    J::set_inf(r);
    if (!J::is_inf(acc)) {
        typename J::cached ca;
        J::to_cached(ca, acc);
        int top = 31;
        while (!((j >> top) & 1u)) top--;
        FT::copy(r.X, acc.X); FT::copy(r.Y, acc.Y); FT::copy(r.Z, acc.Z);
        ECB_NOUNROLL
        for (int bit = top - 1; bit >= 0; bit--) {
            J::dbl(r, r);
            if ((j >> bit) & 1u) J::template add<false>(r, r, ca);
        }
    }
    M::st_pt(bsum + b * M::PW, r);
}

// one level of the tree sum inside each window: slot s += slot s + half, s < half
template <class C>
ECB_DEV void msm_reduce_body(size_t t, u32 NB, u32 half, u32* bsum) {
    typedef Msm<C> M;
    const size_t w = t / half, s = t % half;
    if (s + half >= NB) return;
    u32* pa = bsum + (w * NB + s) * M::PW;
    typename M::J::pt a, bb;
    M::ld_pt(a, pa);
    M::ld_pt(bb, bsum + (w * NB + s + half) * M::PW);
    M::add_pts(a, a, bb);
    M::st_pt(pa, a);
}

// Horner over the windows (slot 0 of each window holds W_w): out = sum_w 2^(c w) W_w, Jacobian, PW words
template <class C>
ECB_DEV void msm_window_body(int c, int nwin, u32 NB, const u32* bsum, u32* out) {
    typedef Msm<C> M;
    typename M::J::pt acc, t;
    M::ld_pt(acc, bsum + (size_t)(nwin - 1) * NB * M::PW);
    for (int w = nwin - 2; w >= 0; w--) {
        for (int s = 0; s < c; s++) M::J::dbl(acc, acc);
        M::ld_pt(t, bsum + (size_t)w * NB * M::PW);
        M::add_pts(acc, acc, t);
    }
    M::st_pt(out, acc);
}

// the partial sums of the devices (np x PW words, gathered on one device) -> one Jacobian point, laid out as the
// X, Y, Z planes of a batch of one for the affine finisher
template <class C>
ECB_DEV void msm_combine_body(int np, const u32* partials, u32* planes) {
    typedef Msm<C> M;
    typename M::J::pt acc, t;
    M::ld_pt(acc, partials);
    for (int i = 1; i < np; i++) {
        M::ld_pt(t, partials + (size_t)i * M::PW);
        M::add_pts(acc, acc, t);
    }
    M::st_pt(planes, acc);
}

}  // namespace ecb
