// kernels2.cuh — GF(2^448-2^224-1) + X448, batched ECDSA verification and batched Ed25519
// verification (per-thread bodies; launchers in eccbatch.cu, host simulation in tests/hostsim/).
#pragma once
#include "kernels.cuh"
#include "sha512.cuh"

namespace ecb {

// =======================================================================================
// GF(p448), p = 2^448 - 2^224 - 1, 14 saturated 32-bit limbs, any 448-bit value ("loose").
// Replaces the 8x56-bit fiat backend src/curve/fiat/p448_solinas_64.rs as used by
// src/curve/curve448.rs:45-122 (field) and :143-174 (pow_const Fermat inverse).
// 2^448 = 2^224 + 1 (mod p): the fold is add-only.  mul = 196 IMAD.WIDE, sqr = 105.
// =======================================================================================
struct fe448 {
    u32 v[14];
};

struct F448 {
    typedef fe448 el;
    static constexpr int N = 14;
    static constexpr bool BLOCK_INV = false;

    ECB_DEV static void set_zero(el& r) {
        ECB_UNROLL
        for (int i = 0; i < 14; i++) r.v[i] = 0;
    }
    ECB_DEV static void set_one(el& r) {
        set_zero(r);
        r.v[0] = 1;
    }
    ECB_DEV static void copy(el& r, const el& a) {
        ECB_UNROLL
        for (int i = 0; i < 14; i++) r.v[i] = a.v[i];
    }
    // r (14 limbs) += c * (2^224 + 1), twice (the second time with the carry of the first)
    ECB_DEV static void fold_top(u32* r, u32 c) {
        ECB_UNROLL
        for (int rep = 0; rep < 2; rep++) {
            r[0] = add_cc(r[0], c);
            ECB_UNROLL
            for (int i = 1; i < 7; i++) r[i] = addc_cc(r[i], 0);
            r[7] = addc_cc(r[7], c);
            ECB_UNROLL
            for (int i = 8; i < 14; i++) r[i] = addc_cc(r[i], 0);
            c = addc(0, 0);
        }
    }
    // reduce a 28-limb product
    ECB_DEV static void reduce(el& r, const u32* t) {
        const u32* lo = t;
        const u32* hi = t + 14;
        const u32* h0 = t + 14;
        const u32* h1 = t + 21;
        u32 A[14], S[7];
        u32 c = add_n<14>(A, lo, hi);  // lo + hi
        // A += h1 (limbs 0..6), propagate
        A[0] = add_cc(A[0], h1[0]);
        ECB_UNROLL
        for (int i = 1; i < 7; i++) A[i] = addc_cc(A[i], h1[i]);
        ECB_UNROLL
        for (int i = 7; i < 14; i++) A[i] = addc_cc(A[i], 0);
        c += addc(0, 0);
        // S = h0 + h1 ; A += S << 224
        c += add_n<7>(S, h0, h1);
        A[7] = add_cc(A[7], S[0]);
        ECB_UNROLL
        for (int i = 1; i < 7; i++) A[7 + i] = addc_cc(A[7 + i], S[i]);
        c += addc(0, 0);
        fold_top(A, c);
        ECB_UNROLL
        for (int i = 0; i < 14; i++) r.v[i] = A[i];
    }
    ECB_DEV static void mul(el& r, const el& a, const el& b) {
        u32 t[28];
        mul_full<14>(t, a.v, b.v);
        reduce(r, t);
    }
    ECB_DEV static void sqr(el& r, const el& a) {
        u32 t[28];
        sqr_full<14>(t, a.v);
        reduce(r, t);
    }
    ECB_DEV static void mul_small(el& r, const el& a, u32 k) {  // k < 2^26
        u32 R[16];
        ECB_UNROLL
        for (int i = 0; i < 16; i++) R[i] = 0;
        mac_chain<7, true>(R, a.v, k);
        mac_chain<7, false>(R + 1, a.v + 1, k);
        fold_top(R, R[14]);
        ECB_UNROLL
        for (int i = 0; i < 14; i++) r.v[i] = R[i];
    }
    ECB_DEV static void add(el& r, const el& a, const el& b) {
        u32 t[14];
        u32 c = add_n<14>(t, a.v, b.v);
        fold_top(t, c);
        ECB_UNROLL
        for (int i = 0; i < 14; i++) r.v[i] = t[i];
    }
    ECB_DEV static void sub(el& r, const el& a, const el& b) {
        u32 t[14];
        u32 c = sub_n<14>(t, a.v, b.v);
        ECB_UNROLL
        for (int rep = 0; rep < 2; rep++) {  // -2^448 = -(2^224 + 1)
            t[0] = sub_cc(t[0], c);
            ECB_UNROLL
            for (int i = 1; i < 7; i++) t[i] = subc_cc(t[i], 0);
            t[7] = subc_cc(t[7], c);
            ECB_UNROLL
            for (int i = 8; i < 14; i++) t[i] = subc_cc(t[i], 0);
            c = subc(0, 0) & 1;
        }
        ECB_UNROLL
        for (int i = 0; i < 14; i++) r.v[i] = t[i];
    }
    // canonical representative: subtract p iff a >= p  <=>  a + 2^224 + 1 carries out of 2^448
    ECB_DEV static void freeze(el& r, const el& a) {
        u32 s[14];
        s[0] = add_cc(a.v[0], 1u);
        ECB_UNROLL
        for (int i = 1; i < 7; i++) s[i] = addc_cc(a.v[i], 0);
        s[7] = addc_cc(a.v[7], 1u);
        ECB_UNROLL
        for (int i = 8; i < 14; i++) s[i] = addc_cc(a.v[i], 0);
        u32 ge = addc(0, 0);
        u32 m = 0u - ge;
        ECB_UNROLL
        for (int i = 0; i < 14; i++) r.v[i] = (s[i] & m) | (a.v[i] & ~m);
    }
    ECB_DEV static u32 is_zero(const el& a) {
        el f;
        freeze(f, a);
        u32 o = 0;
        ECB_UNROLL
        for (int i = 0; i < 14; i++) o |= f.v[i];
        return o == 0 ? 1u : 0u;
    }
    ECB_DEV static void select(el& r, u32 c, const el& a, const el& b) {
        u32 m = 0u - c;
        ECB_UNROLL
        for (int i = 0; i < 14; i++) r.v[i] = (a.v[i] & m) | (b.v[i] & ~m);
    }
    ECB_DEV static void cswap(u32 c, el& a, el& b) {
        u32 m = 0u - c;
        ECB_UNROLL
        for (int i = 0; i < 14; i++) {
            u32 x = (a.v[i] ^ b.v[i]) & m;
            a.v[i] ^= x;
            b.v[i] ^= x;
        }
    }
    ECB_DEV static void sqr_n(el& r, const el& a, int n) {
        sqr(r, a);
        for (int i = 1; i < n; i++) sqr(r, r);
    }
    // safegcd (modinv.cuh); 0 -> 0
    ECB_DEV static void invert(el& r, const el& a) {
        el c;
        freeze(c, a);
        u32 p[14];
        ECB_UNROLL
        for (int i = 0; i < 14; i++) p[i] = (i == 7) ? 0xfffffffeu : 0xffffffffu;
        sg_modinv<14, 15, 37>(r.v, c.v, p);
    }
#ifndef ECB_HOSTSIM
    __device__ __forceinline__ static void invert_warp(el& r, const el& a, const u32* jump) {   // one warp, one element (modinv.cuh)
        el c;
        freeze(c, a);
        u32 p[14];
        ECB_UNROLL
        for (int i = 0; i < 14; i++) p[i] = (i == 7) ? 0xfffffffeu : 0xffffffffu;
        sg_modinv_warp<14, 15, 37>(r.v, c.v, p, jump);
    }
#endif
    // a^(p-2) = a^(2^448 - 2^224 - 3); 0 -> 0
    ECB_DEV static void invert_fermat(el& r, const el& a) {
        el x2, x3, x6, x12, x24, x27, x54, x108, x111, x222, x223, t;
        sqr(t, a);           mul(x2, t, a);
        sqr(t, x2);          mul(x3, t, a);
        sqr_n(t, x3, 3);     mul(x6, t, x3);
        sqr_n(t, x6, 6);     mul(x12, t, x6);
        sqr_n(t, x12, 12);   mul(x24, t, x12);
        sqr_n(t, x24, 3);    mul(x27, t, x3);
        sqr_n(t, x27, 27);   mul(x54, t, x27);
        sqr_n(t, x54, 54);   mul(x108, t, x54);
        sqr_n(t, x108, 3);   mul(x111, t, x3);
        sqr_n(t, x111, 111); mul(x222, t, x111);
        sqr(t, x222);        mul(x223, t, a);
        sqr_n(t, x223, 223); mul(t, t, x222);
        sqr_n(t, t, 2);      mul(r, t, a);
    }
};

// X448: the reference's ladder (src/curve/curve448.rs:263-302) behind protocol::x448::x448
// (src/protocol/x448.rs:34-45): clamp k[0] &= 252, k[55] |= 128; u read as-is (implicitly
// reduced); 448 steps MSB first; result x2 * z2^(p-2) (0 for z2 = 0).
ECB_DEV void x448_body(size_t idx, size_t n, const u32* scalars, const u32* us, u32* planes) {
    u32 k[14];
    fe448 x1, x2, z2, x3, z3;
    ld_words<14>(k, scalars + idx * 14);
    ld_words<14>(x1.v, us + idx * 14);
    k[0] &= 0xfffffffcu;
    k[13] |= 0x80000000u;
    F448::set_one(x2);
    F448::set_zero(z2);
    F448::copy(x3, x1);
    F448::set_one(z3);
    u32 swap = 0;
    for (int wi = 13; wi >= 0; wi--) {
        u32 word = 0;
        ECB_UNROLL
        for (int t = 0; t < 14; t++)
            if (t == wi) word = k[t];
        for (int bi = 31; bi >= 0; bi--) {
            u32 bit = (word >> bi) & 1u;
            swap ^= bit;
            F448::cswap(swap, x2, x3);
            F448::cswap(swap, z2, z3);
            swap = bit;
            fe448 a, aa, b, bb, e, c, d, da, cb, t;
            F448::add(a, x2, z2);
            F448::sqr(aa, a);
            F448::sub(b, x2, z2);
            F448::sqr(bb, b);
            F448::sub(e, aa, bb);
            F448::add(c, x3, z3);
            F448::sub(d, x3, z3);
            F448::mul(da, d, a);
            F448::mul(cb, c, b);
            F448::add(t, da, cb);
            F448::sqr(x3, t);
            F448::sub(t, da, cb);
            F448::sqr(t, t);
            F448::mul(z3, x1, t);
            F448::mul(x2, aa, bb);
            F448::mul_small(t, e, 39082u);
            F448::add(t, bb, t);
            F448::mul(z2, e, t);
        }
    }
    F448::cswap(swap, x2, x3);
    F448::cswap(swap, z2, z3);
    plane_st<14>(planes + 0 * 14 * n, n, idx, x2.v);
    plane_st<14>(planes + 2 * 14 * n, n, idx, z2.v);
}
struct FinX448 {
    const u32* planes; size_t n; u32* out;
    ECB_DEV void pre(size_t idx) const { plane_prefetch<14>(planes, n, idx); }
    ECB_DEV void operator()(size_t idx, const fe448& zinv, u32 zero) const {
        fe448 X, x;
        plane_ld<14>(X.v, planes, n, idx);
        F448::mul(x, X, zinv);
        F448::freeze(x, x);
        u32 m = zero ? 0u : 0xffffffffu;
        ECB_UNROLL
        for (int i = 0; i < 14; i++) x.v[i] &= m;
        st_words<14>(out + idx * 14, x.v);
    }
};

// =======================================================================================
// ECDSA verify_hashed, batched (src/protocol/ecdsa.rs:205-222).
//   prep : decode r || s (Signature::from_bytes :399 — zero or >= n is invalid), s -> Montgomery
//          form in GF(n) into the "Z" plane of `sp` for the scalar-field batch inversion
//   inv  : batch_inv_body<FN> + FinScalarInv  -> s^-1 (Montgomery) in plane 0 of `sp`
//   main : u1 = z s^-1, u2 = r s^-1, R = u2*Q (signed windows of C::WIN bits over a per-thread table, Jacobian) +
//          u1*G (generator comb, no doublings), one general addition; Jacobian result to the point planes
//   fin  : x = X/Z by batch inversion in GF(p); accept iff Z != 0 and x mod n == r
// =======================================================================================
// z = digest_to_scalar(H(message)) before the reduction mod n (src/protocol/ecdsa.rs:288 hash_to_scalar,
// :340 digest_to_scalar = SEC1 bits2int: the leftmost min(8 len, qlen) bits; for p256r1 / p384r1 the
// order fills whole bytes, so that is "left-pad a short digest, keep the first SB bytes of a long one").
// hash: 256 / 384 / 512.  z_out: n x SB bytes big-endian (reduced mod n later, in ecdsa_main_body).
ECB_DEV void ecdsa_hash_z_body(size_t idx, const unsigned char* msgs, const unsigned long long* off, int hash, int SB,
                               unsigned char* z_out) {
    const unsigned char* M = msgs + off[idx];
    size_t mlen = (size_t)(off[idx + 1] - off[idx]);
    unsigned char dg[64];
    int dlen = hash / 8;
    auto at = [&](size_t pos) -> unsigned char { return M[pos]; };
    if (hash == 256) sha256_bytes(dg, mlen, at);
    else sha512_bytes(dg, mlen, at, hash == 384);
    unsigned char* z = z_out + idx * (size_t)SB;
    for (int i = 0; i < SB; i++) {
        int j = dlen <= SB ? i - (SB - dlen) : i;   // left-pad, or keep the leading SB bytes
        z[i] = (j >= 0 && j < dlen) ? dg[j] : 0;
    }
}

template <class C>
ECB_DEV void ecdsa_prep_body(size_t idx, size_t n, const u32* z_be, const u32* rs_be, u32* sp, unsigned char* valid) {
    typedef typename C::FN FN;
    constexpr int NS = FN::N;
    u32 r[NS], s[NS];
    ld_words_be<NS>(r, rs_be + idx * 2 * NS);
    ld_words_be<NS>(s, rs_be + idx * 2 * NS + NS);
    u32 rz = 0, sz = 0;
    ECB_UNROLL
    for (int i = 0; i < NS; i++) { rz |= r[i]; sz |= s[i]; }
    u32 ok = (rz != 0) & (sz != 0) & FN::is_canonical_words(r) & FN::is_canonical_words(s);
    if (!ok) {
        ECB_UNROLL
        for (int i = 0; i < NS; i++) s[i] = 0;
        s[0] = 1;
    }
    typename FN::el sm;
    FN::to_mont(sm, s);
    plane_st<NS>(sp + 2 * (size_t)NS * n, n, idx, sm.v);
    valid[idx] = (unsigned char)ok;
    (void)z_be;
}
template <class C>
struct FinScalarInv {
    typedef typename C::FN FN;
    u32* sp; size_t n;
    ECB_DEV void pre(size_t) const {}
    ECB_DEV void operator()(size_t idx, const typename FN::el& zinv, u32) const { plane_st<FN::N>(sp, n, idx, zinv.v); }
};

template <class C>
ECB_DEV void ecdsa_main_body(size_t idx, size_t n, const u32* q_be, const u32* z_be, const u32* rs_be, const u32* sp,
                             unsigned char* valid, u32* tbl, const u32* gtable, int W, int nwin, u32* planes,
                             unsigned long long* status) {
    typedef Wei<C> WP;
    typedef WeiJ<C> J;
    typedef typename C::F FT;
    typedef typename C::FN FN;
    typedef typename FT::el fe;
    constexpr int N = FT::N;
    constexpr int NS = FN::N;
    constexpr int ES = 5 * N;
    // public key
    u32 xw[N], yw[N];
    ld_words_be<N>(xw, q_be + idx * 2 * N);
    ld_words_be<N>(yw, q_be + idx * 2 * N + N);
    fe qx, qy;
    FT::to_mont(qx, xw);
    FT::to_mont(qy, yw);
    u32 q_ok = 1;
    if (!(FT::is_canonical_words(xw) && FT::is_canonical_words(yw) && WP::on_curve(qx, qy))) {
        report_bad(status, idx, ST_BAD_POINT);
        q_ok = 0;
    }
    // u1 = z * s^-1, u2 = r * s^-1 as plain integers: mont_mul(plain, mont) = plain product
    typename FN::el sinv, zz, rr, u1e, u2e;
    plane_ld<NS>(sinv.v, sp, n, idx);
    ld_words_be<NS>(zz.v, z_be + idx * NS);
    ld_words_be<NS>(rr.v, rs_be + idx * 2 * NS);
    {   // z mod n: 2^(32 NS) < 2n, one conditional subtraction (digest_to_scalar, ecdsa.rs:340)
        u32 d[NS];
        d[0] = sub_cc(zz.v[0], FN::P_::mod(0));
        ECB_UNROLL
        for (int i = 1; i < NS; i++) d[i] = subc_cc(zz.v[i], FN::P_::mod(i));
        u32 borrow = subc(0, 0) & 1;
        ECB_UNROLL
        for (int i = 0; i < NS; i++) zz.v[i] = borrow ? zz.v[i] : d[i];
    }
    typename J::pt acc;
    J::set_inf(acc);
    if (valid[idx] && q_ok) {
        FN::mul(u1e, zz, sinv);
        FN::mul(u2e, rr, sinv);
        u32 u1[NS + 1], u2[NS + 1];
        ECB_UNROLL
        for (int i = 0; i < NS; i++) { u1[i] = u1e.v[i]; u2[i] = u2e.v[i]; }
        u1[NS] = 0;
        u2[NS] = 0;
        // u2*Q: signed windows of C::WIN bits over tbl[0..2^(WIN-1)) = j*Q (Jacobian, cached Z powers)
        wei_build_table<C>(tbl, qx, qy);
        {   // lazy doublings, redone with the checked forms if a fold carried again (kernels.cuh wei_window_loop)
            typename FT::lazy z;
            wei_window_loop<C, FT::LOOSE>(acc, u2, tbl, z);
            if (z.any()) acc = wei_window_loop_checked<C>(u2, tbl);
        }
        // u1*G from the generator comb (Point::mul_base), then one general addition
        typename J::pt accg;
        J::set_inf(accg);
        wei_comb_accumulate<C>(accg, u1, NS + 1, gtable, W, nwin);
        if (!J::is_inf(accg)) {
            typename J::cached cg;
            J::to_cached(cg, accg);
            J::template add<false>(acc, acc, cg);
        }
    }
    plane_st<N>(planes + 0 * (size_t)N * n, n, idx, acc.X.v);
    plane_st<N>(planes + 1 * (size_t)N * n, n, idx, acc.Y.v);
    plane_st<N>(planes + 2 * (size_t)N * n, n, idx, acc.Z.v);
}

template <class C>
struct FinEcdsa {  // x_mod_n(R) == r (ecdsa.rs:382, :218-221); identity => reject
    typedef typename C::F FT;
    typedef typename C::FN FN;
    const u32* planes; size_t n; const u32* rs_be; unsigned char* ok;
    ECB_DEV void pre(size_t idx) const { plane_prefetch<FT::N>(planes, n, idx); }
    ECB_DEV void operator()(size_t idx, const typename FT::el& zinv, u32 zero) const {
        constexpr int N = FT::N;
        constexpr int NS = FN::N;
        typename FT::el X, x;
        plane_ld<N>(X.v, planes, n, idx);
        typename FT::el zi2;
        FT::sqr(zi2, zinv);          // Jacobian: x = X / Z^2
        FT::mul(x, X, zi2);
        u32 xw[N];
        FT::from_mont(xw, x);
        // field_to_scalar (ecdsa.rs:363): FB <= SB here, p < 2n: one conditional subtraction
        u32 d[NS];
        d[0] = sub_cc(xw[0], FN::P_::mod(0));
        ECB_UNROLL
        for (int i = 1; i < NS; i++) d[i] = subc_cc(xw[i], FN::P_::mod(i));
        u32 borrow = subc(0, 0) & 1;
        u32 r[NS];
        ld_words_be<NS>(r, rs_be + idx * 2 * NS);
        u32 diff = 0;
        ECB_UNROLL
        for (int i = 0; i < NS; i++) diff |= (borrow ? xw[i] : d[i]) ^ r[i];
        ok[idx] = (unsigned char)((ok[idx] != 0) & (zero == 0) & (diff == 0));
    }
};

// =======================================================================================
// ECDSA sign_hashed, batched (src/protocol/ecdsa.rs:165-184).  NOT constant-time.
//   prep   : d, k, z decoded; valid = canonical & d != 0 & k != 0; k (or 1) -> Montgomery form in GF(n)
//            into the "Z" plane of `sp` for the scalar-field batch inversion
//   [batch_inv<FN> + FinScalarInv -> k^-1; generator comb on k + affine conversion -> x || y of k G]
//   finish : r = x mod n (field_to_scalar, ecdsa.rs:363); s = k^-1 (z + r d) mod n;
//            valid &= k G finite & r != 0 & s != 0; rs = r || s big-endian (zero when not valid)
// =======================================================================================
// z_any: z is bits2int of a digest (raw-message signing): any SB-byte value, reduced mod n by the
// Montgomery conversion in the finish step; otherwise z is a Scalar and must be canonical
template <class C>
ECB_DEV void ecdsa_sign_prep_body(size_t idx, size_t n, const u32* d_be, const u32* k_be, const u32* z_be, u32* sp, unsigned char* valid,
                                  bool z_any = false) {
    typedef typename C::FN FN;
    constexpr int NS = FN::N;
    u32 d[NS], k[NS], z[NS];
    ld_words_be<NS>(d, d_be + idx * NS);
    ld_words_be<NS>(k, k_be + idx * NS);
    ld_words_be<NS>(z, z_be + idx * NS);
    u32 dz = 0, kz = 0;
    ECB_UNROLL
    for (int i = 0; i < NS; i++) { dz |= d[i]; kz |= k[i]; }
    u32 ok = (dz != 0) & (kz != 0) & FN::is_canonical_words(d) & FN::is_canonical_words(k) & (z_any ? 1u : FN::is_canonical_words(z));
    if (!ok) {
        ECB_UNROLL
        for (int i = 0; i < NS; i++) k[i] = 0;
        k[0] = 1;
    }
    typename FN::el km;
    FN::to_mont(km, k);
    plane_st<NS>(sp + 2 * (size_t)NS * n, n, idx, km.v);
    valid[idx] = (unsigned char)ok;
}
template <class C>
ECB_DEV void ecdsa_sign_finish_body(size_t idx, size_t n, const u32* d_be, const u32* z_be, const u32* kg_xy_be, const unsigned char* kg_inf,
                                    const u32* sp, u32* rs_be, unsigned char* valid) {
    typedef typename C::F FT;
    typedef typename C::FN FN;
    constexpr int N = FT::N;
    constexpr int NS = FN::N;
    u32 xw[N], d[NS], z[NS], r[NS], sw[NS];
    ld_words_be<N>(xw, kg_xy_be + idx * 2 * N);
    ld_words_be<NS>(d, d_be + idx * NS);
    ld_words_be<NS>(z, z_be + idx * NS);
    // field_to_scalar: FB == SB for p256r1 / p384r1 and p < 2n: one conditional subtraction
    u32 t[NS];
    t[0] = sub_cc(xw[0], FN::P_::mod(0));
    ECB_UNROLL
    for (int i = 1; i < NS; i++) t[i] = subc_cc(xw[i], FN::P_::mod(i));
    u32 borrow = subc(0, 0) & 1;
    u32 rz = 0;
    ECB_UNROLL
    for (int i = 0; i < NS; i++) { r[i] = borrow ? xw[i] : t[i]; rz |= r[i]; }
    typename FN::el rm, dm, zm, kinv, acc;
    FN::to_mont(rm, r);
    FN::to_mont(dm, d);
    FN::to_mont(zm, z);
    plane_ld<NS>(kinv.v, sp, n, idx);
    FN::mul(acc, rm, dm);
    FN::add(acc, acc, zm);
    FN::mul(acc, acc, kinv);
    FN::from_mont(sw, acc);
    u32 sz = 0;
    ECB_UNROLL
    for (int i = 0; i < NS; i++) sz |= sw[i];
    u32 ok = (valid[idx] != 0) & (kg_inf[idx] == 0) & (rz != 0) & (sz != 0);
    u32 m = ok ? 0xffffffffu : 0u;
    ECB_UNROLL
    for (int i = 0; i < NS; i++) { r[i] &= m; sw[i] &= m; }
    st_words_be<NS>(rs_be + idx * 2 * NS, r);
    st_words_be<NS>(rs_be + idx * 2 * NS + NS, sw);
    valid[idx] = (unsigned char)ok;
}

// =======================================================================================
// Ed25519 verification, batched (src/protocol/ed25519.rs:119-147) with k = H(R||A||M) mod l
// supplied by the caller.  One kernel, no inversion: the final comparison is projective
// exactly like Point::eq (curve25519.rs:1200).
// =======================================================================================
// decode_point (protocol/ed25519.rs:38-59) + Point::decompress (curve25519.rs:772) +
// sqrt_div (curve25519.rs:258).  Returns 1 and (x, y) on success.
ECB_DEV u32 ed25519_decode(fe25519& x, fe25519& y, const u32* enc) {
    u32 w[8];
    ECB_UNROLL
    for (int i = 0; i < 8; i++) w[i] = enc[i];
    u32 sign = w[7] >> 31;
    w[7] &= 0x7fffffffu;
    u32 ok = F::is_canonical_words(w);
    F::from_words(y, w);
    fe25519 one, u, v, v3, v7, r, chk, t, nu, d, sm1;
    F::set_one(one);
    // (x = 0, sign = 1) <=> y = +-1 with the sign bit set
    {
        fe25519 ny;
        F::neg(ny, one);
        u32 ypm1 = F::eq(y, one) | F::eq(y, ny);
        ok &= (sign & ypm1) ^ 1u;
    }
    F::from_words(d, ED25519_D);
    F::from_words(sm1, ED25519_SQRTM1);
    F::sqr(t, y);
    F::sub(u, t, one);
    F::mul(v, t, d);
    F::add(v, v, one);
    F::sqr(v3, v);
    F::mul(v3, v3, v);
    F::sqr(v7, v3);
    F::mul(v7, v7, v);
    F::mul(t, u, v7);
    F::pow_p58(t, t);
    F::mul(r, u, v3);
    F::mul(r, r, t);
    F::sqr(chk, r);
    F::mul(chk, chk, v);
    F::neg(nu, u);
    u32 correct = F::eq(chk, u);
    u32 flipped = F::eq(chk, nu);
    F::mul(t, r, sm1);
    F::select(r, flipped, t, r);
    ok &= (correct | flipped);
    F::freeze(r, r);
    u32 flip = (r.v[0] & 1u) ^ sign;
    F::neg(t, r);
    F::select(x, flip, t, r);
    return ok;
}

// decode_point / Point::decompress over a batch (protocol/ed25519.rs:38-59, curve25519.rs:772): 32-byte encodings
// -> canonical affine x || y (64 bytes LE) and a presence flag; zero bytes where the reference returns None
// (non-canonical y, no square root, x = 0 with the sign bit set).
ECB_DEV void ed25519_decompress_body(size_t idx, const u32* enc, u32* out_xy, unsigned char* ok) {
    u32 w[8];
    ld_words<8>(w, enc + idx * 8);
    fe25519 x, y;
    u32 good = ed25519_decode(x, y, w);
    F::freeze(x, x);
    F::freeze(y, y);
    u32 m = good ? 0xffffffffu : 0u;
    ECB_UNROLL
    for (int i = 0; i < 8; i++) { x.v[i] &= m; y.v[i] &= m; }
    st_words<8>(out_xy + idx * 16, x.v);
    st_words<8>(out_xy + idx * 16 + 8, y.v);
    ok[idx] = (unsigned char)good;
}

// 64 little-endian bytes -> canonical scalar mod l (reduce_wide_le / init_from_wide_bytes_le)
ECB_DEV void ed25519_reduce_wide(u32* out8, const unsigned char* dg) {
    typedef Mont<ED_FN> FL;
    FL::el lo, hi, r2, one, t;
    ECB_UNROLL
    for (int i = 0; i < 8; i++) {
        lo.v[i] = (u32)dg[4 * i] | ((u32)dg[4 * i + 1] << 8) | ((u32)dg[4 * i + 2] << 16) | ((u32)dg[4 * i + 3] << 24);
        hi.v[i] = (u32)dg[32 + 4 * i] | ((u32)dg[32 + 4 * i + 1] << 8) | ((u32)dg[32 + 4 * i + 2] << 16) | ((u32)dg[32 + 4 * i + 3] << 24);
        r2.v[i] = ED_FN::r2(i);
        one.v[i] = i == 0 ? 1u : 0u;
    }
    FL::mul(hi, hi, r2);       // d_hi * 2^256 mod l
    FL::mul(t, lo, r2);        // d_lo * 2^256 mod l
    FL::mul(lo, t, one);       // d_lo mod l
    FL::add(t, lo, hi);
    ECB_UNROLL
    for (int i = 0; i < 8; i++) out8[i] = t.v[i];
}

// =======================================================================================
// k = reduce_wide_le(SHA-512(R || A || M))  (src/protocol/ed25519.rs:21 reduce_wide_le, :139) and the
// split of the 64-byte signature into R and S: the prologue of verification on raw messages.
//   sig : n x 64 bytes R || S ; a_enc : n x 32 ; msgs + off : concatenated messages, message i is
//   bytes [off[i], off[i+1]) ; outputs r_out, s_out, k_out : n x 32 bytes each
// The 512-bit digest d = d_lo + 2^256 d_hi (little-endian) is reduced with three Montgomery products
// in GF(l): d_hi 2^256 = mont(d_hi, R^2), d_lo = mont(mont(d_lo, R^2), 1) — CIOS accepts any 256-bit
// first operand.
// =======================================================================================
ECB_DEV void ed25519_hash_k_body(size_t idx, const unsigned char* a_enc, const unsigned char* sig, const unsigned char* msgs,
                                 const unsigned long long* off, u32* r_out, u32* s_out, u32* k_out) {
    const unsigned char* R = sig + idx * 64;
    const unsigned char* A = a_enc + idx * 32;
    const unsigned char* M = msgs + off[idx];
    size_t mlen = (size_t)(off[idx + 1] - off[idx]);
    unsigned char dg[64];
    u32 pw[16];                 // R || A as sixteen little-endian words (both 16-byte aligned)
    ld_words<8>(pw, reinterpret_cast<const u32*>(R));
    ld_words<8>(pw + 8, reinterpret_cast<const u32*>(A));
    sha512_prefixed<16>(dg, [&](int i) -> u32 { return pw[i]; }, M, mlen);
    u32 kw[8];
    ed25519_reduce_wide(kw, dg);
    ECB_UNROLL
    for (int i = 0; i < 8; i++) {
        k_out[idx * 8 + i] = kw[i];
        r_out[idx * 8 + i] = (u32)R[4 * i] | ((u32)R[4 * i + 1] << 8) | ((u32)R[4 * i + 2] << 16) | ((u32)R[4 * i + 3] << 24);
        s_out[idx * 8 + i] = (u32)R[32 + 4 * i] | ((u32)R[32 + 4 * i + 1] << 8) | ((u32)R[32 + 4 * i + 2] << 16) | ((u32)R[32 + 4 * i + 3] << 24);
    }
}

// =======================================================================================
// Ed25519 key generation and signing (src/protocol/ed25519.rs:61-110).
// NOT constant-time: secret-dependent table addresses in the comb kernel.
//   expand : h = SHA-512(seed); a = clamp(h[0..32]) mod l (init_from_wide_bytes_le of the clamped
//            value, ed25519.rs:71-74); prefix = h[32..64]
//   nonce  : r = reduce_wide_le(SHA-512(prefix || M))                              (ed25519.rs:96)
//   [comb kernel + FinEdCompressed: R = encode_point(r B), A = encode_point(a B)]
//   finish : k = reduce_wide_le(SHA-512(R || A || M)); S = r + k a mod l; sig = R || S   (:100-109)
// =======================================================================================
ECB_DEV void ed25519_expand_body(size_t idx, const unsigned char* seeds, u32* a_out, u32* prefix_out) {
    const unsigned char* S = seeds + idx * 32;
    unsigned char dg[64], wide[64];
    u32 sw[8];
    ld_words<8>(sw, reinterpret_cast<const u32*>(S));
    sha512_prefixed<8>(dg, [&](int i) -> u32 { return sw[i]; }, nullptr, 0);
    ECB_UNROLL
    for (int i = 0; i < 32; i++) { wide[i] = dg[i]; wide[32 + i] = 0; }
    wide[0] &= 248;
    wide[31] &= 127;
    wide[31] |= 64;
    u32 a[8];
    ed25519_reduce_wide(a, wide);
    ECB_UNROLL
    for (int i = 0; i < 8; i++) {
        a_out[idx * 8 + i] = a[i];
        prefix_out[idx * 8 + i] = (u32)dg[32 + 4 * i] | ((u32)dg[32 + 4 * i + 1] << 8) | ((u32)dg[32 + 4 * i + 2] << 16) | ((u32)dg[32 + 4 * i + 3] << 24);
    }
}
ECB_DEV void ed25519_sign_nonce_body(size_t idx, const u32* prefix, const unsigned char* msgs, const unsigned long long* off, u32* r_out) {
    const u32* P = prefix + idx * 8;
    const unsigned char* M = msgs + off[idx];
    size_t mlen = (size_t)(off[idx + 1] - off[idx]);
    unsigned char dg[64];
    u32 pw[8];
    ld_words<8>(pw, P);
    sha512_prefixed<8>(dg, [&](int i) -> u32 { return pw[i]; }, M, mlen);
    u32 r[8];
    ed25519_reduce_wide(r, dg);
    ECB_UNROLL
    for (int i = 0; i < 8; i++) r_out[idx * 8 + i] = r[i];
}
// sig holds R in bytes [0, 32) of element idx on entry; S is written to bytes [32, 64)
ECB_DEV void ed25519_sign_finish_body(size_t idx, unsigned char* sig, const unsigned char* a_pub, const unsigned char* msgs,
                                      const unsigned long long* off, const u32* a_red, const u32* r_red) {
    typedef Mont<ED_FN> FL;
    const unsigned char* R = sig + idx * 64;
    const unsigned char* A = a_pub + idx * 32;
    const unsigned char* M = msgs + off[idx];
    size_t mlen = (size_t)(off[idx + 1] - off[idx]);
    unsigned char dg[64];
    u32 pw[16];
    ld_words<8>(pw, reinterpret_cast<const u32*>(R));
    ld_words<8>(pw + 8, reinterpret_cast<const u32*>(A));
    sha512_prefixed<16>(dg, [&](int i) -> u32 { return pw[i]; }, M, mlen);
    u32 kw[8];
    ed25519_reduce_wide(kw, dg);
    FL::el k, a, r, r2, one, t;
    ECB_UNROLL
    for (int i = 0; i < 8; i++) {
        k.v[i] = kw[i];
        a.v[i] = a_red[idx * 8 + i];
        r.v[i] = r_red[idx * 8 + i];
        r2.v[i] = ED_FN::r2(i);
        one.v[i] = i == 0 ? 1u : 0u;
    }
    FL::mul(k, k, r2);         // k R
    FL::mul(t, k, a);          // k a   (one Montgomery factor cancels)
    FL::add(t, t, r);          // S = r + k a mod l
    FL::canon(t, t);
    u32* out = reinterpret_cast<u32*>(sig + idx * 64 + 32);
    ECB_UNROLL
    for (int i = 0; i < 8; i++) out[i] = t.v[i];
}

ECB_DEV void ed25519_verify_body(size_t idx, size_t n, const u32* a_enc, const u32* s_le, const u32* k_le,
                                 const u32* table, int W, int nwin, int stride, u32* tbl, u32* planes, unsigned char* out_ok) {
    u32 aw[8], s[8], k[8];
    ld_words<8>(aw, a_enc + idx * 8);
    ld_words<8>(s, s_le + idx * 8);
    ld_words<8>(k, k_le + idx * 8);
    fe25519 ax, ay;
    u32 ok = ed25519_decode(ax, ay, aw);
    ok &= lt_words8(s, ED25519_L);
    ok &= lt_words8(k, ED25519_L);
    ge_p3 acc;
    ge_identity(acc);
    if (ok) {
        // [S]B from the comb
        const u32 half = 1u << (W - 1);
        ge_p3 sb;
        ge_identity(sb);
        for (int i = 0; i < nwin; i++) {
            u32 neg;
            u32 d = booth_digit(s, 8, W, i, neg);
            if (d != 0) {
                ge_niels e;
                const u32* src = table + ((size_t)i * half + (d - 1)) * (size_t)stride;
                ld_words<8>(e.yp.v, src);
                ld_words<8>(e.ym.v, src + 8);
                ld_words<8>(e.t2d.v, src + 16);
                ge_niels_cneg(e, neg);
                ge_madd<true>(sb, sb, e);
            }
        }
        // [k](-A): signed 4-bit windows over 8 cached multiples
        ge_p3 P;
        {
            fe25519 nax;
            F::neg(nax, ax);
            ge_from_affine(P, nax, ay);
            ge_cached c, c1;
            ge_to_cached(c1, P);
            acc = P;
            ECB_NOUNROLL
            for (int j = 1; j <= 8; j++) {
                ge_to_cached(c, acc);
                u32* d = tbl + (j - 1) * 32;
                st_words<8>(d + 0, c.yp.v); st_words<8>(d + 8, c.ym.v); st_words<8>(d + 16, c.Z.v); st_words<8>(d + 24, c.t2d.v);
                if (j < 8) ge_add_cached<true>(acc, acc, c1);
            }
        }
        ge_identity(acc);
        ECB_NOUNROLL
        for (int i = 63; i >= 0; i--) {
            if (i != 63) {
                ECB_NOUNROLL
                for (int r = 0; r < 4; r++) ge_double_rt<true>(acc, acc, r == 3 ? 1u : 0u);
            }
            u32 neg;
            u32 d = booth_digit(k, 8, 4, i, neg);
            if (d != 0) {
                ge_cached c;
                const u32* sp = tbl + (d - 1) * 32;
                ld_words_rw<8>(c.yp.v, sp); ld_words_rw<8>(c.ym.v, sp + 8); ld_words_rw<8>(c.Z.v, sp + 16); ld_words_rw<8>(c.t2d.v, sp + 24);
                ge_cached_cneg(c, neg);
                ge_add_cached<true, true>(acc, acc, c);
            }
        }
        // lhs = [S]B + [k](-A)
        ge_cached cs;
        ge_to_cached(cs, sb);
        ge_add_cached<false>(acc, acc, cs);
    }
    out_ok[idx] = (unsigned char)ok;
    plane_st<8>(planes + 0 * 8 * n, n, idx, acc.X.v);
    plane_st<8>(planes + 1 * 8 * n, n, idx, acc.Y.v);
    plane_st<8>(planes + 2 * 8 * n, n, idx, acc.Z.v);
}
// lhs == R  <=>  encode_point(lhs) == the signature's R bytes: decode_point (protocol/ed25519.rs:38)
// accepts exactly the canonical encodings of curve points and encode_point is its inverse, so a
// non-canonical / off-curve / (x = 0, sign = 1) R can never equal an encoding and is rejected just as
// the reference's decode step rejects it — without paying a second square-root exponentiation.
struct FinEdVerify {
    const u32* planes; size_t n; const u32* r_enc; unsigned char* ok;
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
        u32 rw[8], diff = 0;
        ld_words<8>(rw, r_enc + idx * 8);
        ECB_UNROLL
        for (int i = 0; i < 8; i++) diff |= rw[i] ^ y.v[i];
        ok[idx] = (unsigned char)((ok[idx] != 0) & (diff == 0));
    }
};

}  // namespace ecb
