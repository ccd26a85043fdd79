// weier.cuh — short-Weierstrass curves of the batch path (p256r1, p384r1, BLS12-381 G1): curve
// descriptions, the on-curve check and Jacobian point arithmetic.
//
// The reference's generic projective::Point<FE> uses homogeneous coordinates with the complete
// Renes–Costello–Batina formulas (src/curve/projective.rs:340 add_different_am3, :586 double_am3,
// :268 add_different_a0, :544 double_a0; INFINITY = (0:1:0) :152).  The kernels compute the same
// group law in Jacobian coordinates (see WeiJ below): only canonical affine results are observable.
#pragma once
#include "mont_kinds.cuh"

namespace ecb {

template <class FT>
struct wpoint {
    typename FT::el X, Y, Z;
};

// Curve description: field F, scalar field FN, A_M3, b()/b3() in the Montgomery domain,
// byte sizes on the wire (big-endian).
struct CurveP256 {
    typedef Mont<P256_FP> F;
    typedef Mont<P256_FN> FN;
    static constexpr bool A_M3 = true;
    static constexpr int FB = 32, SB = 32;          // field / scalar bytes
    static constexpr bool COFACTOR = false;      // #E(Fp) != group order: small-order points exist
    static constexpr int SBITS = 256;
#ifndef ECB_P256_WIN
#define ECB_P256_WIN 5      // measured on B200, n = 2^20: 4 -> 5 bits = 65 -> 52 additions for 8 more table entries: +2.4 %
#endif
    static constexpr int WIN = ECB_P256_WIN;        // signed windows of the variable-base kernels (table of 2^(WIN-1) multiples per thread)
    ECB_DEV static u32 b(int i) { return P256_B[i]; }
    ECB_DEV static u32 b3(int i) { return P256_B3[i]; }
    ECB_DEV static u32 gx(int i) { return P256_GX[i]; }
    ECB_DEV static u32 gy(int i) { return P256_GY[i]; }
    ECB_DEV static u32 sqrt_e(int i) { return P256_SQRT_E[i]; }
};
struct CurveP384 {
    typedef Mont<P384_FP> F;
    typedef Mont<P384_FN> FN;
    static constexpr bool A_M3 = true;
    static constexpr int FB = 48, SB = 48;
    static constexpr bool COFACTOR = false;      // #E(Fp) != group order: small-order points exist
    static constexpr int SBITS = 384;
#ifndef ECB_P384_WIN
#define ECB_P384_WIN 5      // +4.4 % (n = 2^18)
#endif
    static constexpr int WIN = ECB_P384_WIN;
    ECB_DEV static u32 b(int i) { return P384_B[i]; }
    ECB_DEV static u32 b3(int i) { return P384_B3[i]; }
    ECB_DEV static u32 gx(int i) { return P384_GX[i]; }
    ECB_DEV static u32 gy(int i) { return P384_GY[i]; }
    ECB_DEV static u32 sqrt_e(int i) { return P384_SQRT_E[i]; }
};
struct CurveBLSG1 {
    typedef Mont<BLS_FP> F;
    typedef Mont<BLS_FR> FN;
    static constexpr bool A_M3 = false;
    static constexpr int FB = 48, SB = 32;
    static constexpr bool COFACTOR = true;      // #E(Fp) != group order: small-order points exist
    static constexpr int SBITS = 255;
#ifndef ECB_BLS_WIN
#define ECB_BLS_WIN 5       // +2.3 %
#endif
    static constexpr int WIN = ECB_BLS_WIN;
    ECB_DEV static u32 b(int i) { return BLSG1_B[i]; }
    ECB_DEV static u32 b3(int i) { return BLSG1_B3[i]; }
    ECB_DEV static u32 gx(int i) { return BLSG1_GX[i]; }
    ECB_DEV static u32 gy(int i) { return BLSG1_GY[i]; }
    ECB_DEV static u32 sqrt_e(int i) { return BLSG1_SQRT_E[i]; }
    ECB_DEV static u32 beta(int i) { return BLSG1_BETA[i]; }
};

// secp256k1 (the reference's p256k1, src/curve/sec2/p256k1.rs): a = 0, b = 7 — the same a = 0 path as BLS12-381 G1
// (projective.rs:268, :544, :842, :945), generic CIOS Montgomery fields on 8 limbs
struct CurveK256 {
    typedef Mont<K256_FP> F;
    typedef Mont<K256_FN> FN;
    static constexpr bool A_M3 = false;
    static constexpr int FB = 32, SB = 32;
    static constexpr bool COFACTOR = false;      // #E(Fp) != group order: small-order points exist
    static constexpr int SBITS = 256;
#ifndef ECB_K256_WIN
#define ECB_K256_WIN 5      // +2.8 %
#endif
    static constexpr int WIN = ECB_K256_WIN;
    ECB_DEV static u32 b(int i) { return K256_B[i]; }
    ECB_DEV static u32 b3(int i) { return K256_B3[i]; }
    ECB_DEV static u32 gx(int i) { return K256_GX[i]; }
    ECB_DEV static u32 gy(int i) { return K256_GY[i]; }
    ECB_DEV static u32 sqrt_e(int i) { return K256_SQRT_E[i]; }
};

template <class C>
struct Wei {
    typedef typename C::F F;
    typedef typename F::el fe;
    typedef wpoint<F> pt;
    static constexpr int N = F::N;

    ECB_DEV static void get_b(fe& r) {
        ECB_UNROLL
        for (int i = 0; i < N; i++) r.v[i] = C::b(i);
    }
    ECB_DEV static void get_b3(fe& r) {
        ECB_UNROLL
        for (int i = 0; i < N; i++) r.v[i] = C::b3(i);
    }
    ECB_DEV static void set_inf(pt& r) {
        F::set_zero(r.X);
        F::set_one(r.Y);
        F::set_zero(r.Z);
    }
    // y^2 == x^3 + a x + b  (affine::Point::from_coordinate_ct, src/curve/affine.rs:88)
    ECB_DEV static u32 on_curve(const fe& x, const fe& y) {
        fe l, r, t, b;
        F::sqr(l, y);
        F::sqr(r, x);
        F::mul(r, r, x);
        if (C::A_M3) {
            F::dbl(t, x);
            F::add(t, t, x);
            F::sub(r, r, t);
        }
        get_b(b);
        F::add(r, r, b);
        return F::eq(l, r);
    }

};

// =======================================================================================
// Jacobian coordinates (X : Y : Z), x = X/Z^2, y = Y/Z^3, infinity <=> Z = 0.
//
// The batch kernels use these instead of the complete projective formulas above: only the
// canonical affine result is observable, the group law is the same, and a doubling
// costs 3M + 5S (a = -3) / 2M + 5S (a = 0) instead of 8M + 3S + 2 m_b — fewer passes through the
// integer-multiply pipe, which is what bounds these kernels.  The addition formulas are not
// complete, so the exceptional inputs (P = Q, P = -Q, either operand at infinity) are detected and
// handled explicitly; the kernels are variable-time anyway (public data).
// =======================================================================================
template <class C>
struct WeiJ {
    typedef typename C::F F;
    typedef typename F::el fe;
    static constexpr int N = F::N;
    static constexpr int TBL = 1 << (C::WIN - 1);   // entries of a per-thread table of multiples
    struct pt {
        fe X, Y, Z;
    };
    // table entry: (X, Y, Z, Z^2, Z^3); AFF entries have Z = 1 and only X, Y are meaningful
    struct cached {
        fe X, Y, Z, ZZ, ZZZ;
    };

    ECB_DEV static void set_inf(pt& r) {
        F::set_one(r.X);
        F::set_one(r.Y);
        F::set_zero(r.Z);
    }
    ECB_DEV static u32 is_inf(const pt& p) { return F::is_zero(p.Z); }

    // dbl-2001-b (a = -3): 3M + 5S ; dbl-2009-l (a = 0): 2M + 5S.  Z = 0 stays Z = 0.
    // LZ: the field additions only COUNT a carry out of their fold (mont.cuh lazy forms); dbl() checks the counters once.
    template <bool LZ>
    ECB_DEV static void dbl_impl(pt& r, const pt& p, typename F::lazy& z) {
        if (C::A_M3) {
            fe delta, gamma, beta, alpha, t0, t1, X3, Y3, Z3;
            F::sqr_ni(delta, p.Z);
            F::sqr_ni(gamma, p.Y);
            F::mul_ni(beta, p.X, gamma);
            F::template sub_z<LZ>(t0, p.X, delta, z);
            F::template add_z<LZ>(t1, p.X, delta, z);
            F::mul_ni(t0, t0, t1);
            F::template mul_small_z<3, LZ>(alpha, t0, z);   // 3 (X - delta)(X + delta)
            F::template add_z<LZ>(t1, p.Y, p.Z, z);
            F::sqr_ni(Z3, t1);
            F::template sub2_z<LZ>(Z3, Z3, gamma, delta, z);   // (Y + Z)^2 - gamma - delta
            F::sqr_ni(X3, alpha);
            F::template mul_small_z<4, LZ>(t0, beta, z);    // 4 beta
            F::template sub2_z<LZ>(X3, X3, t0, t0, z);      // alpha^2 - 8 beta
            F::template sub_z<LZ>(t0, t0, X3, z);
            F::mul_ni(Y3, alpha, t0);
            F::sqr_ni(t1, gamma);
            F::template mul_small_z<8, LZ>(t1, t1, z);      // 8 gamma^2
            F::template sub_z<LZ>(Y3, Y3, t1, z);
            F::copy(r.X, X3);
            F::copy(r.Y, Y3);
            F::copy(r.Z, Z3);
        } else {
            fe A, B, Cc, D, E, Fv, t, X3, Y3, Z3;
            F::sqr_ni(A, p.X);
            F::sqr_ni(B, p.Y);
            F::sqr_ni(Cc, B);
            F::template add_z<LZ>(t, p.X, B, z);
            F::sqr_ni(D, t);
            F::template sub2_z<LZ>(D, D, A, Cc, z);
            F::template mul_small_z<2, LZ>(D, D, z);        // 2((X + B)^2 - A - C)
            F::template mul_small_z<3, LZ>(E, A, z);        // 3A
            F::sqr_ni(Fv, E);
            F::template sub2_z<LZ>(X3, Fv, D, D, z);        // F - 2D
            F::mul_ni(Z3, p.Y, p.Z);
            F::template mul_small_z<2, LZ>(Z3, Z3, z);
            F::template sub_z<LZ>(t, D, X3, z);
            F::mul_ni(Y3, E, t);
            F::template mul_small_z<8, LZ>(Cc, Cc, z);      // 8C
            F::template sub_z<LZ>(Y3, Y3, Cc, z);
            F::copy(r.X, X3);
            F::copy(r.Y, Y3);
            F::copy(r.Z, Z3);
        }
    }
    ECB_DEV static void dbl(pt& r, const pt& p) {   // every fold checked
        typename F::lazy z;
        dbl_impl<false>(r, p, z);
    }
    // LZ = true: folds only counted in z — the CALLER must look at z.any() and redo its work with LZ = false if set
    template <bool LZ>
    ECB_DEV static void dbl_z(pt& r, const pt& p, typename F::lazy& z) { dbl_impl<LZ>(r, p, z); }
    ECB_DEV static void to_cached(cached& c, const pt& p) {
        F::copy(c.X, p.X);
        F::copy(c.Y, p.Y);
        F::copy(c.Z, p.Z);
        F::sqr_ni(c.ZZ, p.Z);
        F::mul_ni(c.ZZZ, c.ZZ, p.Z);
    }
    ECB_DEV static void cached_from_affine(cached& c, const fe& x, const fe& y) {
        F::copy(c.X, x);
        F::copy(c.Y, y);
        F::set_one(c.Z);
        F::set_one(c.ZZ);
        F::set_one(c.ZZZ);
    }
    // r = p + q, q a finite table entry (never infinity).  add-2007-bl with the Z2 powers cached:
    // 10M + 4S; QAFF (Z2 = 1): 7M + 4S (madd-2007-bl).
    template <bool QAFF>
    ECB_DEV static void add(pt& r, const pt& p, const cached& q) {
        if (F::is_zero(p.Z)) {  // infinity + q
            F::copy(r.X, q.X);
            F::copy(r.Y, q.Y);
            if (QAFF) F::set_one(r.Z); else F::copy(r.Z, q.Z);
            return;
        }
        if (!QAFF && C::COFACTOR && F::is_zero(q.Z)) {   // p + O (see add_mem)
            F::copy(r.X, p.X); F::copy(r.Y, p.Y); F::copy(r.Z, p.Z);
            return;
        }
        fe Z1Z1, U1, U2, S1, S2, H, I, J, rr, V, t, X3, Y3, Z3;
        F::sqr_ni(Z1Z1, p.Z);
        if (QAFF) F::copy(U1, p.X); else F::mul_ni(U1, p.X, q.ZZ);
        F::mul_ni(U2, q.X, Z1Z1);
        if (QAFF) F::copy(S1, p.Y); else F::mul_ni(S1, p.Y, q.ZZZ);
        F::mul_ni(t, p.Z, Z1Z1);
        F::mul_ni(S2, q.Y, t);
        F::sub(H, U2, U1);
        F::sub(rr, S2, S1);
        if (F::is_zero(H)) {
            if (F::is_zero(rr)) {  // p == q: double
                dbl(r, p);
            } else {               // p == -q
                set_inf(r);
            }
            return;
        }
        fe HH;
        F::dbl(rr, rr);             // r = 2 (S2 - S1)
        F::sqr_ni(HH, H);
        F::template mul_small<4>(I, HH);   // I = 4 H^2
        F::mul_ni(J, H, I);
        F::mul_ni(V, U1, I);
        F::sqr_ni(X3, rr);
        F::sub_2x(X3, X3, J, V);    // r^2 - J - 2V
        F::sub(t, V, X3);
        F::mul_ni(Y3, rr, t);
        F::mul_ni(t, S1, J);
        F::sub2(Y3, Y3, t, t);      // r (V - X3) - 2 S1 J
        if (QAFF) {
            F::add(t, p.Z, H);      // Z3 = (Z1 + H)^2 - Z1Z1 - HH = 2 Z1 H
            F::sqr_ni(Z3, t);
            F::sub2(Z3, Z3, Z1Z1, HH);
        } else {
            F::add(t, p.Z, q.Z);
            F::sqr_ni(Z3, t);
            F::sub2(Z3, Z3, Z1Z1, q.ZZ);
            F::mul_ni(Z3, Z3, H);
        }
        F::copy(r.X, X3);
        F::copy(r.Y, Y3);
        F::copy(r.Z, Z3);
    }
    // r = p + (neg ? -q : q) with q = the cached entry stored at `e` (5N words: X, Y, Z, Z^2, Z^3).
    // Same formulas as add<false>; the entry's fields are loaded where they are consumed so that
    // only one of them is live at a time (the whole entry would cost 5N registers).
    // ENDO (BLS12-381 G1 only): the entry is read as phi(q) = (beta X2, Y2, Z2): one product more (GLV, kernels3.cuh).
    template <class LD, bool ENDO = false>
    ECB_DEV static void add_mem(pt& r, const pt& p, const u32* e, u32 neg, LD ld) {
        fe q;
        if (F::is_zero(p.Z)) {  // infinity + q
            ld(r.X.v, e);
            if constexpr (ENDO) {
                ECB_UNROLL
                for (int i = 0; i < N; i++) q.v[i] = C::beta(i);
                F::mul_ni(r.X, r.X, q);
            }
            ld(q.v, e + N);
            F::neg(r.Y, q);
            F::select(r.Y, neg, r.Y, q);
            ld(r.Z.v, e + 2 * N);
            return;
        }
        if (C::COFACTOR) {
            // On a curve with cofactor > 1 a table entry j * P is the identity when ord(P) divides j (BLS12-381
            // E(Fp) has points of order 3, and Point::mul accepts any curve point, g1.rs:375): p + O = p.
            // The formulas below would give Z3 = 0 instead.  Prime-order curves never store such an entry.
            ld(q.v, e + 2 * N);
            if (F::is_zero(q)) {
                F::copy(r.X, p.X); F::copy(r.Y, p.Y); F::copy(r.Z, p.Z);
                return;
            }
        }
        fe Z1Z1, U1, U2, S1, S2, H, I, J, rr, V, t, X3, Y3, Z3, HH;
        F::sqr_ni(Z1Z1, p.Z);
        ld(q.v, e + 3 * N);                 // Z2^2
        F::mul_ni(U1, p.X, q);
        ld(q.v, e);                         // X2
        if constexpr (ENDO) {
            ECB_UNROLL
            for (int i = 0; i < N; i++) U2.v[i] = C::beta(i);
            F::mul_ni(q, q, U2);
        }
        F::mul_ni(U2, q, Z1Z1);
        F::sub(H, U2, U1);
        ld(q.v, e + 4 * N);                 // Z2^3
        F::mul_ni(S1, p.Y, q);
        F::mul_ni(t, p.Z, Z1Z1);
        ld(q.v, e + N);                     // Y2
        F::mul_ni(S2, q, t);
        F::neg(t, S2);
        F::select(S2, neg, t, S2);
        F::sub(rr, S2, S1);
        if (F::is_zero(H)) {
            if (F::is_zero(rr)) {
                dbl(r, p);
            } else {
                set_inf(r);
            }
            return;
        }
        F::dbl(rr, rr);
        F::sqr_ni(HH, H);
        F::template mul_small<4>(I, HH);
        F::mul_ni(J, H, I);
        F::mul_ni(V, U1, I);
        F::sqr_ni(X3, rr);
        F::sub_2x(X3, X3, J, V);
        F::sub(t, V, X3);
        F::mul_ni(Y3, rr, t);
        F::mul_ni(t, S1, J);
        F::sub2(Y3, Y3, t, t);
        ld(q.v, e + 2 * N);                 // Z2
        F::add(t, p.Z, q);
        F::sqr_ni(Z3, t);
        ld(q.v, e + 3 * N);                 // Z2^2 again
        F::sub2(Z3, Z3, Z1Z1, q);
        F::mul_ni(Z3, Z3, H);
        F::copy(r.X, X3);
        F::copy(r.Y, Y3);
        F::copy(r.Z, Z3);
    }
    ECB_DEV static void cached_cneg(cached& c, u32 neg) {
        fe ny;
        F::neg(ny, c.Y);
        F::select(c.Y, neg, ny, c.Y);
    }
};

}  // namespace ecb
