// kernels3.cuh — wire formats either side of the Weierstrass path:
//   PointAffine::decompress (src/curve/affine.rs:48, fiat/curve_macros.rs:221) for p256r1 / p384r1 /
//   bls12_381 G1, and the BLS12-381 G1 standard encodings (src/curve/bls12_381/serialize.rs) with the
//   prime-order-subgroup test of g1.rs:105.  One thread per element, no batch inversion needed.
#pragma once
#include "kernels.cuh"

namespace ecb {

// r = a^((p+1)/4) — the square-root candidate for p = 3 mod 4 (p256r1.rs:68, p384r1.rs:71,
// bls12_381/fp.rs:64; the reference walks per-prime addition chains, any chain gives the same
// element) — by 2-bit fixed windows over the exponent; returns [r^2 == a].
template <class C>
ECB_DEV u32 wei_sqrt(typename C::F::el& r, const typename C::F::el& a) {
    typedef typename C::F FT;
    typedef typename FT::el fe;
    constexpr int N = FT::N;
    fe a2, a3, acc, chk;
    FT::sqr_ni(a2, a);
    FT::mul_ni(a3, a2, a);
    FT::set_one(acc);
    ECB_NOUNROLL
    for (int i = 16 * N - 1; i >= 0; i--) {
        FT::sqr_ni(acc, acc);
        FT::sqr_ni(acc, acc);
        u32 d = (C::sqrt_e(i >> 4) >> (2 * (i & 15))) & 3u;
        if (d == 1) FT::mul_ni(acc, acc, a);
        else if (d == 2) FT::mul_ni(acc, acc, a2);
        else if (d == 3) FT::mul_ni(acc, acc, a3);
    }
    FT::sqr_ni(chk, acc);
    r = acc;
    return FT::eq(chk, a);
}

// rhs = x^3 + a x + b
template <class C>
ECB_DEV void wei_rhs(typename C::F::el& r, const typename C::F::el& x) {
    typedef typename C::F FT;
    typename FT::el t, b;
    FT::sqr_ni(r, x);
    FT::mul_ni(r, r, x);
    if (C::A_M3) {
        FT::dbl(t, x);
        FT::add(t, t, x);
        FT::sub(r, r, t);
    }
    Wei<C>::get_b(b);
    FT::add(r, r, b);
}

// PointAffine::decompress: x (FB bytes BE, canonical) + sign (0 = Sign::Positive: even y, 1 = Negative:
// odd y; FieldElement::sign reads bit 0 of the canonical value, field_macros.rs:557) -> x || y, ok.
// Not ok (output zero): x >= p (FieldElement::from_bytes gives None) or x^3 + a x + b not a square.
template <class C>
ECB_DEV void wei_decompress_body(size_t idx, const u32* x_be, const unsigned char* sign, u32* out_xy, unsigned char* ok) {
    typedef typename C::F FT;
    typedef typename FT::el fe;
    constexpr int N = FT::N;
    u32 xw[N], yw[N];
    ld_words_be<N>(xw, x_be + idx * N);
    u32 good = FT::is_canonical_words(xw);
    fe x, rhs, y;
    FT::to_mont(x, xw);
    wei_rhs<C>(rhs, x);
    good &= wei_sqrt<C>(y, rhs);
    FT::from_mont(yw, y);
    if ((yw[0] & 1u) != (sign[idx] ? 1u : 0u)) {
        FT::neg(y, y);
        FT::from_mont(yw, y);
    }
    u32 m = good ? 0xffffffffu : 0u;
    ECB_UNROLL
    for (int i = 0; i < N; i++) { xw[i] &= m; yw[i] &= m; }
    st_words_be<N>(out_xy + idx * 2 * N, xw);
    st_words_be<N>(out_xy + idx * 2 * N + N, yw);
    ok[idx] = (unsigned char)good;
}

// ---- BLS12-381 G1 ---------------------------------------------------------------------------
// a > b on little-endian limbs
template <int N>
ECB_DEV u32 gt_words(const u32* a, const u32* b) {
    (void)sub_cc(b[0], a[0]);
    ECB_UNROLL
    for (int i = 1; i < N; i++) (void)subc_cc(b[i], a[i]);
    return subc(0u, 0u) & 1u;  // borrow of b - a
}

// [|x|] P for the curve seed |x| = 0xd201000000010000 (g1.rs:68 mul_by_abs_x): 63 doublings and 5
// additions of the cached base.
template <bool QAFF>
ECB_DEV void bls_mul_by_abs_x(WeiJ<CurveBLSG1>::pt& acc, const WeiJ<CurveBLSG1>::cached& base) {
    typedef WeiJ<CurveBLSG1> J;
    const unsigned long long X = 0xd201000000010000ull;
    ECB_NOUNROLL
    for (int i = 62; i >= 0; i--) {
        J::dbl(acc, acc);
        if ((X >> i) & 1ull) J::template add<QAFF>(acc, acc, base);
    }
}

// PointAffine::is_in_subgroup (g1.rs:105): sigma(P) = (beta x, y) must equal -[x^2]P.
ECB_DEV u32 bls_g1_in_subgroup(const fe_mont<12>& x, const fe_mont<12>& y) {
    typedef CurveBLSG1 C;
    typedef C::F FT;
    typedef WeiJ<C> J;
    typedef FT::el fe;
    J::cached c;
    J::pt acc;
    J::cached_from_affine(c, x, y);
    acc.X = x;
    acc.Y = y;
    FT::set_one(acc.Z);
    bls_mul_by_abs_x<true>(acc, c);
    if (J::is_inf(acc)) return 0;          // [|x|]P = infinity: sigma(P) is affine, -[x^2]P is not
    J::to_cached(c, acc);
    bls_mul_by_abs_x<false>(acc, c);
    if (J::is_inf(acc)) return 0;
    // (beta x, y) == (X / Z^2, -Y / Z^3)
    fe beta, zz, zzz, l, ny;
    ECB_UNROLL
    for (int i = 0; i < 12; i++) beta.v[i] = C::beta(i);
    FT::sqr_ni(zz, acc.Z);
    FT::mul_ni(zzz, zz, acc.Z);
    FT::mul_ni(l, x, beta);
    FT::mul_ni(l, l, zz);
    u32 same = FT::eq(l, acc.X);
    FT::mul_ni(l, y, zzz);
    FT::neg(ny, acc.Y);
    same &= FT::eq(l, ny);
    return same;
}

// PointAffine::from_compressed (check != 0) / from_compressed_oncurve_only (check == 0),
// serialize.rs:286-321 with read_compressed_flags :117 and read_compressed_affine :172.
// ok = 0 and zero output for: compression flag clear, the identity (no affine point), x >= p,
// x^3 + 4 not a square, and (check) a point outside the prime-order subgroup.
ECB_DEV void bls_g1_from_compressed_body(size_t idx, const u32* enc, int check, u32* out_xy, unsigned char* ok) {
    typedef CurveBLSG1 C;
    typedef C::F FT;
    typedef FT::el fe;
    constexpr int N = 12;
    u32 xw[N], yw[N], nw[N];
    ld_words_be<N>(xw, enc + idx * N);
    const u32 flags = xw[N - 1] >> 29;     // bit 2: compressed, bit 1: infinity, bit 0: y is the larger root
    xw[N - 1] &= 0x1fffffffu;
    u32 good = ((flags & 4u) ? 1u : 0u) & ((flags & 2u) ? 0u : 1u);
    good &= FT::is_canonical_words(xw);
    fe x, rhs, y, ny;
    FT::to_mont(x, xw);
    wei_rhs<C>(rhs, x);
    good &= wei_sqrt<C>(y, rhs);
    FT::neg(ny, y);
    FT::from_mont(yw, y);
    FT::from_mont(nw, ny);
    const u32 largest = gt_words<N>(yw, nw);   // y > p - y  <=>  y > (p - 1) / 2   (Fp::is_largest, serialize.rs:139)
    if (largest != (flags & 1u)) {
        y = ny;
        ECB_UNROLL
        for (int i = 0; i < N; i++) yw[i] = nw[i];
    }
    if (good && check) good = bls_g1_in_subgroup(x, y);
    u32 m = good ? 0xffffffffu : 0u;
    ECB_UNROLL
    for (int i = 0; i < N; i++) { xw[i] &= m; yw[i] &= m; }
    st_words_be<N>(out_xy + idx * 2 * N, xw);
    st_words_be<N>(out_xy + idx * 2 * N + N, yw);
    ok[idx] = (unsigned char)good;
}

// PointAffine::from_uncompressed / from_uncompressed_oncurve_only (serialize.rs:330-380; flags :129-140,
// coordinates :207-221): 96 bytes x || y big-endian, the three flag bits in the top of byte 0.  The compression
// and sort bits must be clear; both coordinates canonical; on the curve; in G1 when `check`.  The identity
// encoding (0x40 then zeros) is valid as an encoding but PointAffine cannot hold it: ok = 0 as the reference
// returns None, and inf (optional) = 1 so that a caller working with projective points can tell it apart.
ECB_DEV void bls_g1_from_uncompressed_body(size_t idx, const u32* enc, int check, u32* out_xy, unsigned char* inf, unsigned char* ok) {
    typedef CurveBLSG1 C;
    typedef C::F FT;
    typedef FT::el fe;
    constexpr int N = 12;
    u32 xw[N], yw[N];
    ld_words_be<N>(xw, enc + idx * 2 * N);
    ld_words_be<N>(yw, enc + idx * 2 * N + N);
    const u32 flags = xw[N - 1] >> 29;     // bit 2: compressed, bit 1: infinity, bit 0: sort
    xw[N - 1] &= 0x1fffffffu;
    u32 payload = 0;
    ECB_UNROLL
    for (int i = 0; i < N; i++) payload |= xw[i] | yw[i];
    const u32 is_inf = ((flags & 5u) == 0u && (flags & 2u) && payload == 0u) ? 1u : 0u;
    u32 good = (flags == 0u) ? 1u : 0u;
    good &= FT::is_canonical_words(xw) & FT::is_canonical_words(yw);
    fe x, y;
    FT::to_mont(x, xw);
    FT::to_mont(y, yw);
    good &= Wei<C>::on_curve(x, y);
    if (good && check) good = bls_g1_in_subgroup(x, y);
    u32 m = good ? 0xffffffffu : 0u;
    ECB_UNROLL
    for (int i = 0; i < N; i++) { xw[i] &= m; yw[i] &= m; }
    st_words_be<N>(out_xy + idx * 2 * N, xw);
    st_words_be<N>(out_xy + idx * 2 * N + N, yw);
    if (inf) inf[idx] = (unsigned char)is_inf;
    ok[idx] = (unsigned char)good;
}
// Point::to_uncompressed (serialize.rs:412-): x || y as given, or 0x40 followed by zeros for the identity
ECB_DEV void bls_g1_to_uncompressed_body(size_t idx, const u32* xy, const unsigned char* inf, u32* enc) {
    constexpr int N = 12;
    u32 w[2 * N];
    ld_words<2 * N>(w, xy + idx * 2 * N);
    if (inf && inf[idx]) {
        ECB_UNROLL
        for (int i = 0; i < 2 * N; i++) w[i] = 0;
        w[0] = 0x40u;   // byte 0 of the big-endian encoding is the low byte of the first little-endian word
    }
    st_words<2 * N>(enc + idx * 2 * N, w);
}

// Point::to_compressed (serialize.rs:400-420): x with the compression flag and the sort flag; the
// identity (inf[idx] != 0) is 0xc0 followed by zeros.  Coordinates are taken as given (canonical).
ECB_DEV void bls_g1_to_compressed_body(size_t idx, const u32* xy, const unsigned char* inf, u32* enc) {
    constexpr int N = 12;
    u32 xw[N], yw[N], nw[N];
    ld_words_be<N>(xw, xy + idx * 2 * N);
    ld_words_be<N>(yw, xy + idx * 2 * N + N);
    nw[0] = sub_cc(BLS_FP::mod(0), yw[0]);
    ECB_UNROLL
    for (int i = 1; i < N; i++) nw[i] = subc_cc(BLS_FP::mod(i), yw[i]);
    (void)subc(0u, 0u);
    u32 largest = gt_words<N>(yw, nw);
    xw[N - 1] |= 0x80000000u | (largest << 29);
    if (inf && inf[idx]) {
        ECB_UNROLL
        for (int i = 0; i < N; i++) xw[i] = 0;
        xw[N - 1] = 0xc0000000u;
    }
    st_words_be<N>(enc + idx * N, xw);
}

// =======================================================================================
// BLS12-381 G1, inputs PROMISED to lie in the prime-order subgroup (option bls12_381_g1_glv): k P through the
// endomorphism phi(x, y) = (beta x, y) = [-x^2] P (the map of the reference's subgroup test, g1.rs:55, :105).
//     k = q x^2 + rem  (0 <= rem < x^2 < 2^128, q < 2^128)   =>   k P = rem P + q (x^2 P) = rem P - q phi(P),
// two 128-bit scalars over ONE table of multiples (phi of an entry costs one product): 25 x 5 doublings instead of
// 51 x 5.  Same result as wei_mul_body for every point of G1 — and only for those: phi is not a multiplication on
// the cofactor part of E(Fp), which Point::mul (g1.rs:375) accepts, so this is never the default.
// =======================================================================================
// k (8 words, little-endian, below the group order) -> k1 = rem, k2 = q, 5 words each with a zero top word.
// x^2 = 2^32 y, so q = (k >> 32) div y and rem = ((k >> 32) mod y) 2^32 + (k mod 2^32): a 224-by-96-bit restoring division.
ECB_DEV void bls_glv_split(u32* k1, u32* k2, const u32* k) {
    const u64 yl = ((u64)BLSG1_XSQ[2] << 32) | BLSG1_XSQ[1], yh = BLSG1_XSQ[3];
    u64 rl = 0, rh = 0;
    u32 q[7];
    ECB_UNROLL
    for (int w = 7; w >= 1; w--) {
        u32 cur = k[w], qw = 0;
        ECB_NOUNROLL
        for (int b = 0; b < 32; b++) {
            rh = (rh << 1) | (rl >> 63);
            rl = (rl << 1) | (cur >> 31);
            cur <<= 1;
            const u32 ge = (rh > yh || (rh == yh && rl >= yl)) ? 1u : 0u;
            if (ge) {
                rh = rh - yh - (rl < yl ? 1u : 0u);
                rl = rl - yl;
            }
            qw = (qw << 1) | ge;
        }
        q[w - 1] = qw;
    }
    k1[0] = k[0]; k1[1] = (u32)rl; k1[2] = (u32)(rl >> 32); k1[3] = (u32)rh; k1[4] = 0;
    k2[0] = q[0]; k2[1] = q[1]; k2[2] = q[2]; k2[3] = q[3]; k2[4] = 0;   // q[4..6] = 0 for k < 2^256 / 2^127
}
template <class C>
ECB_DEV void wei_mul_glv_body(size_t idx, size_t n, const u32* scalars, const u32* points, const unsigned char* inf_in,
                              u32* tbl, u32* planes, unsigned long long* status) {
    typedef WeiJ<C> J;
    typedef Wei<C> W;
    typedef typename C::F FT;
    typedef typename C::FN FNT;
    typedef typename FT::el fe;
    constexpr int N = FT::N;
    constexpr int NS = C::SB / 4;
    constexpr int ES = 5 * N;
    static_assert(NS == 8, "256-bit scalars");
    u32 k[NS];
    ld_words_be<NS>(k, scalars + idx * NS);
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
        u32 k1[5], k2[5];
        bls_glv_split(k1, k2, k);
        wei_build_table<C>(tbl, px, py);
        constexpr int WIN = C::WIN;
        constexpr int NWIN = (128 + 1 + WIN - 1) / WIN;
        auto ld = [](u32* dst, const u32* src) { ld_words_rw<N>(dst, src); };
        ECB_NOUNROLL
        for (int i = NWIN - 1; i >= 0; i--) {
            if (i != NWIN - 1) {
                ECB_NOUNROLL
                for (int r = 0; r < WIN; r++) J::dbl(acc, acc);
            }
            u32 neg;
            u32 d = booth_digit(k1, 5, WIN, i, neg);
            if (d != 0) J::add_mem(acc, acc, tbl + (d - 1) * ES, neg, ld);
            d = booth_digit(k2, 5, WIN, i, neg);                       // - q phi(P): the opposite sign
            if (d != 0) J::template add_mem<decltype(ld), true>(acc, acc, tbl + (d - 1) * ES, neg ^ 1u, ld);
        }
    }
    plane_st<N>(planes + 0 * (size_t)N * n, n, idx, acc.X.v);
    plane_st<N>(planes + 1 * (size_t)N * n, n, idx, acc.Y.v);
    plane_st<N>(planes + 2 * (size_t)N * n, n, idx, acc.Z.v);
}

}  // namespace ecb
