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

// ---- bucket sums, balanced over the SORTED index array ------------------------------------------------------
// The counting sort leaves idx[0 .. L) ordered by bucket.  A thread per bucket would make the running time follow the
// LARGEST bucket: the top window of a 256-bit scalar holds one or two bits, so half of all points share one bucket
// (a 2^18 batch on p256k1 took 0.79 s that way), and equal scalars put everything into one bucket per window.  So the
// array is cut into segments of S entries, one thread each: every thread does S mixed additions whatever the
// distribution.  A bucket that lies inside one segment is summed there and written to bsum[b]; a bucket cut by
// segment boundaries leaves pieces — per segment at most a head piece (its first run: part[2t]) and a tail piece (its
// last run: part[2t + 1]) — which msm_finish_body adds up.
ECB_DEV size_t msm_bucket_of(const u32* offs, size_t nb, size_t pos) {   // last b with offs[b] <= pos
    size_t lo = 0, hi = nb;                                              // invariant: offs[lo] <= pos < offs[hi] (offs[nb] = L)
    while (hi - lo > 1) {
        size_t mid = (lo + hi) >> 1;
        if ((size_t)offs[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
}
template <class C>
ECB_DEV void msm_segment_body(size_t t, size_t S, size_t nb, const u32* offs, const u32* hist, const u32* idx, const u32* pm, u32* bsum,
                              u32* part) {
    typedef Msm<C> M;
    typedef typename M::J J;
    typedef typename C::F FT;
    constexpr int N = FT::N;
    const size_t L = (size_t)offs[nb - 1] + hist[nb - 1];
    size_t pos = t * S;
    if (pos >= L) return;
    const size_t hi = pos + S < L ? pos + S : L;
    // ONE loop over the segment's entries, the same trip count in every lane, so that the warp stays converged on the
    // addition (the expensive part); the end of a bucket is a short divergent side step.  (Nested per-bucket loops let
    // the lanes drift apart: the additions then ran with partial masks and a 2^20 batch took 30 % longer.)
    size_t b = msm_bucket_of(offs, nb, pos);
    size_t bs = offs[b], be = bs + hist[b], run_start = pos;
    bool first = true;
    typename J::pt acc;
    J::set_inf(acc);
    ECB_NOUNROLL
    for (size_t e = pos; e < hi; e++) {
        if (e == be) {                                   // bucket b ends here: it is whole if this run began at its start
            M::st_pt(run_start == bs ? bsum + b * M::PW : part + 2 * t * M::PW, acc);
            do b++; while (hist[b] == 0);                // empty buckets share their offset with the next one
            bs = e;
            be = bs + hist[b];
            run_start = e;
            first = false;
            J::set_inf(acc);
        }
        const u32 v = idx[e];
        const u32* src = pm + (size_t)(v & 0x7fffffffu) * 2 * N;
        typename J::cached q;
        ld_words<N>(q.X.v, src);
        ld_words<N>(q.Y.v, src + N);
        J::cached_cneg(q, v >> 31);
        J::template add<true>(acc, acc, q);
    }
    if (run_start == bs && hi == be) M::st_pt(bsum + b * M::PW, acc);      // the whole bucket
    else M::st_pt(part + (2 * t + (first ? 0 : 1)) * M::PW, acc);          // a piece: the head of this segment, or its tail
}
// the pieces of bucket b, if it was cut: segment tf holds its beginning (as that segment's tail, or as its head when the
// bucket starts exactly on the boundary), every later segment up to tl holds a head piece.  Pieces t in [lo, hi)
// with stride `step` (one thread: lo = tf + 1, step = 1; a warp: lane-strided) are added to acc.
template <class C>
ECB_DEV void msm_pieces(typename WeiJ<C>::pt& acc, const u32* part, size_t lo, size_t hi, size_t step) {
    typedef Msm<C> M;
    typename M::J::pt p;
    ECB_NOUNROLL
    for (size_t t = lo; t < hi; t += step) {
        M::ld_pt(p, part + 2 * t * M::PW);
        M::add_pts(acc, acc, p);
    }
}
// bucket b: B = its sum (unscaled) -> bsum[b].  Returns through (tf, tl) the segments it spans; the caller adds the
// middle pieces of a HEAVY bucket (tl - tf > heavy) with a whole warp and passes them in `mid`; lighter ones are added here.
template <class C>
ECB_DEV void msm_finish_body(size_t b, size_t S, const u32* offs, const u32* hist, const u32* part, u32* bsum, const typename WeiJ<C>::pt* mid) {
    typedef Msm<C> M;
    typedef typename M::J J;
    const u32 cnt = hist[b];
    typename J::pt acc;
    if (cnt == 0) {
        J::set_inf(acc);
        M::st_pt(bsum + b * M::PW, acc);
        return;
    }
    const size_t bs = offs[b], be = bs + cnt, tf = bs / S, tl = (be - 1) / S;
    if (tf == tl) return;                                   // summed inside one segment: bsum[b] is already there
    M::ld_pt(acc, part + (2 * tf + (bs == tf * S ? 0 : 1)) * M::PW);
    if (mid) M::add_pts(acc, acc, *mid);
    else msm_pieces<C>(acc, part, tf + 1, tl + 1, 1);
    M::st_pt(bsum + b * M::PW, acc);
}
// ---- sum_j j * B_j by running sums, CH buckets per thread ----------------------------------------------------
// Chunk q of window w covers the buckets j = a + 1 .. a + CH (a = q CH):
//     sum_j j B_j = sum_i (a + i) B_(a+i) = a * T + sum_i i B_(a+i),   T = sum_i B_(a+i).
// Walking i = CH .. 1 with run += B, acc += run gives acc = sum_i i B_(a+i) and run = T: two additions per bucket
// instead of a 15-bit double-and-add (232 field products), and ONE a * T per chunk.  The result replaces the chunk's
// first slot; the tree sum then runs over the NB / CH chunk results of each window.
template <class C>
ECB_DEV void msm_chunk_body(size_t t, u32 NB, u32 CH, u32* bsum, u32* csum) {
    typedef Msm<C> M;
    typedef typename M::J J;
    typedef typename C::F FT;
    const u32 NC = NB / CH;
    const size_t w = t / NC;
    const u32 q = (u32)(t % NC), a = q * CH;
    const u32* base = bsum + (w * NB + a) * M::PW;
    typename J::pt run, acc, p;
    J::set_inf(run);
    J::set_inf(acc);
    ECB_NOUNROLL
    for (int i = (int)CH - 1; i >= 0; i--) {
        M::ld_pt(p, base + (size_t)i * M::PW);
        M::add_pts(run, run, p);
        M::add_pts(acc, acc, run);
    }
    if (a != 0 && !J::is_inf(run)) {                         // acc += a * T
        typename J::cached ct;
        J::to_cached(ct, run);
        int top = 31;
        while (!((a >> top) & 1u)) top--;
        typename J::pt r;
        FT::copy(r.X, run.X); FT::copy(r.Y, run.Y); FT::copy(r.Z, run.Z);
        ECB_NOUNROLL
        for (int bit = top - 1; bit >= 0; bit--) {
            J::dbl(r, r);
            if ((a >> bit) & 1u) J::template add<false>(r, r, ct);
        }
        M::add_pts(acc, acc, r);
    }
    M::st_pt(csum + t * M::PW, acc);
}

// one level of the tree sum inside each window: slot s += slot s + half, s < half  (NB slots per window)
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
