// mont.cuh — generic N x 32-bit Montgomery field (R = 2^(32N)), canonical values in [0, p).
//
// Replaces (for the batch path) the reference's word-by-word Montgomery fiat backends
// src/curve/fiat/p256_64.rs:265 (mul) :562 (square) :859 (add) :941 (sub) :1013 (opp)
// :1084 (from_montgomery) :1259 (to_montgomery), p384_64.rs, bls12_381_64.rs and the
// *_scalar_64.rs twins, as wired by fiat_field_montgomery_impl!
// (src/curve/fiat/field_macros.rs:511-770).  Elements stay canonical (< p) in the
// Montgomery domain exactly like the fiat code (file header NOTE of p256_64.rs).
//
// mul is CIOS on the even/odd accumulator layout of limb.cuh: per row two chains for
// a*b_i and two for m*p, 2N^2 IMAD.WIDE.U32 total, one right-shift by a limb per row
// done by renaming registers.
#pragma once
#include "limb.cuh"

namespace ecb {

template <int N_>
struct fe_mont {
    u32 v[N_];
};

// P supplies: N, INV (= -p^-1 mod 2^32) and accessors mod(i), r1(i) (= R mod p, "one"),
// r2(i) (= R^2 mod p) over little-endian 32-bit limbs held in `__constant__ const`
// arrays; with the loops unrolled the limbs fold to immediates.
template <class P>
struct Mont {
    static constexpr int N = P::N;
    typedef P P_;
    typedef fe_mont<P::N> el;

    ECB_DEV static void set_zero(el& r) {
        ECB_UNROLL
        for (int i = 0; i < N; i++) r.v[i] = 0;
    }
    ECB_DEV static void set_one(el& r) {
        ECB_UNROLL
        for (int i = 0; i < N; i++) r.v[i] = P::r1(i);
    }
    ECB_DEV static void copy(el& r, const el& a) {
        ECB_UNROLL
        for (int i = 0; i < N; i++) r.v[i] = a.v[i];
    }

    // r = t - p if t >= p (t has an extra top bit `hi`), else t
    ECB_DEV static void final_sub(el& r, const u32* t, u32 hi) {
        u32 d[N];
        d[0] = sub_cc(t[0], P::mod(0));
        ECB_UNROLL
        for (int i = 1; i < N; i++) d[i] = subc_cc(t[i], P::mod(i));
        u32 borrow = subc(0, 0) & 1;
        // take d when hi==1 or no borrow
        u32 take = hi | (borrow ^ 1u);
        u32 m = 0u - take;
        ECB_UNROLL
        for (int i = 0; i < N; i++) r.v[i] = (d[i] & m) | (t[i] & ~m);
    }

    ECB_DEV static void mul(el& r, const el& a, const el& b) {
        // ev[k] sits on (relative) limb k, od[k] on limb k+1; L is a single word on limb 0
        u32 ev[N + 2], od[N + 2];
        ECB_UNROLL
        for (int i = 0; i < N + 2; i++) { ev[i] = 0; od[i] = 0; }
        u32 L = 0;
        u32 pm[N];
        ECB_UNROLL
        for (int i = 0; i < N; i++) pm[i] = P::mod(i);
        ECB_UNROLL
        for (int i = 0; i < N; i++) {
            u32 bi = b.v[i];
            // limb 0: ev[0] += L ; its carry enters the odd chain (limb 1)
            ev[0] = add_cc(ev[0], L);
            od[0] = madc_lo_cc(a.v[1], bi, od[0]);
            od[1] = madc_hi_cc(a.v[1], bi, od[1]);
            ECB_UNROLL
            for (int k = 1; k < N / 2; k++) {
                od[2 * k] = madc_lo_cc(a.v[2 * k + 1], bi, od[2 * k]);
                od[2 * k + 1] = madc_hi_cc(a.v[2 * k + 1], bi, od[2 * k + 1]);
            }
            od[N] = addc(od[N], 0);
            mac_chain<N / 2, true>(ev, a.v, bi);
            u32 m = ev[0] * P::INV;
            mac_chain<N / 2, true>(od, pm + 1, m);
            mac_chain<N / 2, true>(ev, pm, m);
            // now ev[0] == 0 ; divide by 2^32: od becomes the even array, ev>>2 words the odd one
            L = ev[1];
            u32 nev[N + 2], nod[N + 2];
            ECB_UNROLL
            for (int k = 0; k < N + 2; k++) nev[k] = od[k];
            ECB_UNROLL
            for (int k = 0; k < N; k++) nod[k] = ev[k + 2];
            nod[N] = 0;
            nod[N + 1] = 0;
            ECB_UNROLL
            for (int k = 0; k < N + 2; k++) { ev[k] = nev[k]; od[k] = nod[k]; }
        }
        // t = ev + (od << 32) + L, N+1 limbs (t < 2p)
        u32 t[N + 1];
        t[0] = add_cc(ev[0], L);
        ECB_UNROLL
        for (int k = 1; k <= N; k++) t[k] = addc_cc(ev[k], od[k - 1]);
        final_sub(r, t, t[N]);
    }
    ECB_DEV static void sqr(el& r, const el& a) { mul(r, a, a); }

    ECB_DEV static void add(el& r, const el& a, const el& b) {
        u32 t[N];
        u32 c = add_n<N>(t, a.v, b.v);
        final_sub(r, t, c);
    }
    ECB_DEV static void sub(el& r, const el& a, const el& b) {
        u32 t[N];
        u32 bw = sub_n<N>(t, a.v, b.v);
        u32 m = 0u - bw;
        r.v[0] = add_cc(t[0], P::mod(0) & m);
        ECB_UNROLL
        for (int i = 1; i < N; i++) r.v[i] = addc_cc(t[i], P::mod(i) & m);
    }
    ECB_DEV static void neg(el& r, const el& a) {
        el z;
        set_zero(z);
        sub(r, z, a);
    }
    ECB_DEV static void dbl(el& r, const el& a) { add(r, a, a); }

    ECB_DEV static u32 is_zero(const el& a) {
        u32 o = 0;
        ECB_UNROLL
        for (int i = 0; i < N; i++) o |= a.v[i];
        return o == 0 ? 1u : 0u;
    }
    ECB_DEV static u32 eq(const el& a, const el& b) {
        u32 o = 0;
        ECB_UNROLL
        for (int i = 0; i < N; i++) o |= a.v[i] ^ b.v[i];
        return o == 0 ? 1u : 0u;
    }
    ECB_DEV static void select(el& r, u32 c, const el& a, const el& b) {
        u32 m = 0u - c;
        ECB_UNROLL
        for (int i = 0; i < N; i++) r.v[i] = (a.v[i] & m) | (b.v[i] & ~m);
    }

    // plain integer (canonical, little-endian limbs) <-> Montgomery domain
    ECB_DEV static void to_mont(el& r, const u32* w) {
        el t, r2;
        ECB_UNROLL
        for (int i = 0; i < N; i++) { t.v[i] = w[i]; r2.v[i] = P::r2(i); }
        mul(r, t, r2);
    }
    ECB_DEV static void from_mont(u32* w, const el& a) {
        el one, t;
        set_zero(one);
        one.v[0] = 1;
        mul(t, a, one);
        ECB_UNROLL
        for (int i = 0; i < N; i++) w[i] = t.v[i];
    }
    // w < p ?
    ECB_DEV static u32 is_canonical_words(const u32* w) {
        u32 d = sub_cc(w[0], P::mod(0));
        ECB_UNROLL
        for (int i = 1; i < N; i++) d = subc_cc(w[i], P::mod(i));
        (void)d;
        return subc(0, 0) & 1;  // borrow <=> w < p
    }

    // limb i of p-2 (borrow propagated from the low limbs)
    ECB_DEV static u32 pm2_limb(int i) {
        if (i == 0) return P::mod(0) - 2u;
        u32 borrow = P::mod(0) < 2u ? 1u : 0u;
        for (int j = 1; j < i; j++) borrow = (borrow && P::mod(j) == 0u) ? 1u : 0u;
        return P::mod(i) - borrow;
    }

    // a^(p-2) by square-and-multiply over the bits of p-2 (MSB first); 0 -> 0.
    // The result is the field inverse, so any correct chain gives the same bits as
    // the reference's Fermat chains (p256r1.rs:49-65, p384r1.rs:50-69) and safegcd
    // (bls12_381/fp.rs:55-57).
    ECB_DEV static void invert(el& r, const el& a) {
        el acc;
        set_one(acc);
        for (int i = N - 1; i >= 0; i--) {
            u32 e = pm2_limb(i);
            for (int bit = 31; bit >= 0; bit--) {
                sqr(acc, acc);
                if ((e >> bit) & 1) mul(acc, acc, a);
            }
        }
        copy(r, acc);
    }
};

}  // namespace ecb
