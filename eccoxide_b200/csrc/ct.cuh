// ct.cuh — constant-time forms of the fixed-base scalar multiplications, for SECRET scalars (key generation
// and signing).
//
// The reference's secret-dependent paths are constant-time: Point::mul_base reads EVERY entry of a window's
// table and keeps one by masks (select_from_table, curve25519.rs:862-869; projective.rs:427), its inverses
// are fixed Fermat chains (curve25519.rs:155-200).  The large-batch kernels of this library are not — they
// index a multi-gigabyte comb by the scalar's digits and run a variable-time safegcd.  The kernels here
// restore the reference's discipline on the GPU:
//   * a small comb, signed 4-bit windows: 8 entries per window, the whole table (64 x 8 x 96 B = 48 KB for
//     edwards25519) staged into SHARED memory once per block;
//   * every thread reads all 8 entries of every window at addresses that do not depend on any scalar (all
//     lanes read the same word: a broadcast), and keeps the one it needs with masks; digit 0 and the sign
//     are masks as well; no branch, no address and no loop count depends on a secret;
//   * the affine conversion is Montgomery's trick around the fixed 265-step Fermat chain
//     (batch_inv_body<.., CT = true>).
// What stays public: the batch size, the validity of the inputs (a non-canonical scalar is reported, as the
// reference's Scalar::from_bytes returns None), and the order of the windows.
#pragma once
#include "kernels.cuh"

namespace ecb {

// all-ones when a == b
ECB_DEV u32 ct_eq_mask(u32 a, u32 b) {
    u32 x = a ^ b;
    return (u32)((((unsigned long long)x) - 1ull) >> 63) * 0xffffffffu;   // x == 0 -> borrow -> 1
}
// signed radix-16 digit of the low 5 bits of view (Booth): |d| in [0, 8], neg = 1 for a negative digit
ECB_DEV u32 ct_booth4(u32 view, u32& neg) {
    view &= 0x1fu;
    u32 s = view >> 4;
    u32 d = (view + 1u) >> 1;
    u32 dn = 16u - d;
    u32 r = d ^ ((d ^ dn) & (0u - s));
    neg = s & ~(ct_eq_mask(r, 0u) & 1u);
    return r;
}

#define ECB_CT_W 4
#define ECB_CT_ED_NWIN 64                       // ceil(254 / 4): covers 256 bits, as the reference's 64 nibbles do
#define ECB_CT_ED_WORDS (ECB_CT_ED_NWIN * 8 * 24)  // 12288 words = 48 KB

// acc = k * B for a canonical scalar k (< l < 2^253), all table reads from `tbl` (shared memory),
// entry (i, j) = j * 16^i * B in niels form at tbl + (i * 8 + j - 1) * 24, j = 1..8.
ECB_DEV void ed25519_mul_base_ct(ge_p3& acc, const u32* k, const u32* tbl) {
    u32 v[8];
    booth_reg_init<8>(v, k);
    ge_identity(acc);
    ECB_NOUNROLL
    for (int i = 0; i < ECB_CT_ED_NWIN; i++) {
        u32 neg;
        const u32 d = ct_booth4(v[0], neg);
        booth_reg_shift<8>(v, ECB_CT_W);
        ge_niels e;
        ECB_UNROLL
        for (int w = 0; w < 8; w++) { e.yp.v[w] = 0; e.ym.v[w] = 0; e.t2d.v[w] = 0; }
        const u32* row = tbl + (size_t)i * 8 * 24;
        ECB_UNROLL
        for (u32 j = 1; j <= 8; j++) {
            const u32 m = ct_eq_mask(d, j);
            const u32* src = row + (j - 1) * 24;
            ECB_UNROLL
            for (int w = 0; w < 8; w++) {
                e.yp.v[w] |= src[w] & m;
                e.ym.v[w] |= src[8 + w] & m;
                e.t2d.v[w] |= src[16 + w] & m;
            }
        }
        const u32 z = ct_eq_mask(d, 0u) & 1u;     // digit 0: the identity (1, 1, 0)
        e.yp.v[0] |= z;
        e.ym.v[0] |= z;
        ge_niels_cneg(e, neg);
        ge_madd_rt(acc, acc, e, i + 1 < ECB_CT_ED_NWIN);   // the window index is public
    }
}

// one thread per scalar; tbl_g: the W = 4 comb in global memory, copied to shared memory by the block
template <bool CLAMP>
ECB_DEV void ed25519_mul_base_ct_body(size_t idx, size_t n, const u32* scalars, const u32* tbl, u32* planes, unsigned long long* status) {
    u32 k[9];
    ed25519_load_scalar<CLAMP>(k, idx, scalars, status);   // canonical check: validity is public
    ge_p3 acc;
    ed25519_mul_base_ct(acc, k, tbl);
    plane_st<8>(planes + 0 * 8 * n, n, idx, acc.X.v);
    plane_st<8>(planes + 1 * 8 * n, n, idx, acc.Y.v);
    plane_st<8>(planes + 2 * 8 * n, n, idx, acc.Z.v);
}

// ---- short Weierstrass, a = -3 (p256r1, p384r1): constant-time k * G for ECDSA nonces --------------------------
// The variable-time comb adds in Jacobian coordinates and branches on the exceptional cases (P = +-Q, infinity).
// Here the accumulator is homogeneous projective (X : Y : Z) and every window uses the COMPLETE addition of
// Renes-Costello-Batina (eprint 2015/1060, Algorithm 4: 12 M + 2 m_b) — the formulas the reference itself
// runs (projective.rs:340-423) — so no input needs a branch: digit 0 adds (0 : 1 : 0).  The table is the W = 4
// comb of the curve (entries (x, y) in the Montgomery domain, 8 per window) in shared memory, scanned with masks.
template <class C>
ECB_DEV void wei_add_complete_am3(typename C::F::el& X3, typename C::F::el& Y3, typename C::F::el& Z3, const typename C::F::el& X1,
                                  const typename C::F::el& Y1, const typename C::F::el& Z1, const typename C::F::el& X2,
                                  const typename C::F::el& Y2, const typename C::F::el& Z2) {
    typedef typename C::F F;
    typename F::el t0, t1, t2, t3, t4, b, x3, y3, z3;
    Wei<C>::get_b(b);
    F::mul_ni(t0, X1, X2); F::mul_ni(t1, Y1, Y2); F::mul_ni(t2, Z1, Z2);
    F::add_ct(t3, X1, Y1); F::add_ct(t4, X2, Y2); F::mul_ni(t3, t3, t4);
    F::add_ct(t4, t0, t1); F::sub_ct(t3, t3, t4); F::add_ct(t4, Y1, Z1);
    F::add_ct(x3, Y2, Z2); F::mul_ni(t4, t4, x3); F::add_ct(x3, t1, t2);
    F::sub_ct(t4, t4, x3); F::add_ct(x3, X1, Z1); F::add_ct(y3, X2, Z2);
    F::mul_ni(x3, x3, y3); F::add_ct(y3, t0, t2); F::sub_ct(y3, x3, y3);
    F::mul_ni(z3, b, t2); F::sub_ct(x3, y3, z3); F::add_ct(z3, x3, x3);
    F::add_ct(x3, x3, z3); F::sub_ct(z3, t1, x3); F::add_ct(x3, t1, x3);
    F::mul_ni(y3, b, y3); F::add_ct(t1, t2, t2); F::add_ct(t2, t1, t2);
    F::sub_ct(y3, y3, t2); F::sub_ct(y3, y3, t0); F::add_ct(t1, y3, y3);
    F::add_ct(y3, t1, y3); F::add_ct(t1, t0, t0); F::add_ct(t0, t1, t0);
    F::sub_ct(t0, t0, t2); F::mul_ni(t1, t4, y3); F::mul_ni(t2, t0, y3);
    F::mul_ni(y3, x3, z3); F::add_ct(y3, y3, t2); F::mul_ni(x3, t3, x3);
    F::sub_ct(x3, x3, t1); F::mul_ni(z3, t4, z3); F::mul_ni(t1, t3, t0);
    F::add_ct(z3, z3, t1);
    F::copy(X3, x3); F::copy(Y3, y3); F::copy(Z3, z3);
}

template <class C>
struct WeiCt {
    static constexpr int N = C::F::N;
    static constexpr int NS = C::SB / 4;
    static constexpr int NWIN = (C::SBITS + 1 + ECB_CT_W - 1) / ECB_CT_W;   // 65 (p256r1), 97 (p384r1)
    static constexpr int WORDS = NWIN * 8 * 2 * N;                          // 33 KB / 74.5 KB of shared memory
};

// planes <- k * G in the Jacobian form the finishers expect (x = X / Z^2, y = Y / Z^3; Z = 0 at infinity)
template <class C>
ECB_DEV void wei_mul_base_ct_body(size_t idx, size_t n, const u32* scalars, const u32* tbl, u32* planes, unsigned long long* status) {
    typedef typename C::F F;
    typedef typename C::FN FNT;
    typedef typename F::el fe;
    constexpr int N = WeiCt<C>::N, NS = WeiCt<C>::NS, NV = NS + 1;
    u32 k[NS + 1];
    ld_words_be<NS>(k, scalars + idx * NS);
    k[NS] = 0;
    if (!FNT::is_canonical_words(k)) {          // validity is public (Scalar::from_bytes -> None)
        report_bad(status, idx, ST_NONCANONICAL_SCALAR);
        ECB_UNROLL
        for (int i = 0; i < NS; i++) k[i] = 0;
    }
    u32 v[NV];
    booth_reg_init<NV>(v, k);
    fe X, Y, Z, one;
    F::set_one(one);
    F::set_zero(X);
    F::set_one(Y);
    F::set_zero(Z);
    ECB_NOUNROLL
    for (int i = 0; i < WeiCt<C>::NWIN; i++) {
        u32 neg;
        const u32 d = ct_booth4(v[0], neg);
        booth_reg_shift<NV>(v, ECB_CT_W);
        fe ex, ey, ez, ny;
        ECB_UNROLL
        for (int w = 0; w < N; w++) { ex.v[w] = 0; ey.v[w] = 0; }
        const u32* row = tbl + (size_t)i * 8 * 2 * N;
        ECB_UNROLL
        for (u32 j = 1; j <= 8; j++) {
            const u32 m = ct_eq_mask(d, j);
            const u32* src = row + (j - 1) * 2 * N;
            ECB_UNROLL
            for (int w = 0; w < N; w++) {
                ex.v[w] |= src[w] & m;
                ey.v[w] |= src[N + w] & m;
            }
        }
        const u32 z = ct_eq_mask(d, 0u) & 1u;
        F::neg_ct(ny, ey);
        F::select(ey, neg, ny, ey);
        F::select(ey, z, one, ey);               // digit 0: (0 : 1 : 0)
        F::set_zero(ez);
        F::select(ez, z, ez, one);
        wei_add_complete_am3<C>(X, Y, Z, X, Y, Z, ex, ey, ez);
    }
    fe zz, xj, yj;
    F::sqr_ni(zz, Z);
    F::mul_ni(xj, X, Z);
    F::mul_ni(yj, Y, zz);
    plane_st<N>(planes + 0 * (size_t)N * n, n, idx, xj.v);
    plane_st<N>(planes + 1 * (size_t)N * n, n, idx, yj.v);
    plane_st<N>(planes + 2 * (size_t)N * n, n, idx, Z.v);
}

}  // namespace ecb
