// weier.cuh — short-Weierstrass homogeneous projective arithmetic with the complete
// Renes–Costello–Batina formulas (eprint 2015/1060), a = -3 (Alg. 4/5/6) and a = 0 (Alg. 7/8/9).
//
// Same group law as the reference's generic projective::Point<FE>
// (src/curve/projective.rs:340 add_different_am3, :586 double_am3, :268 add_different_a0,
// :544 double_a0); INFINITY = (0:1:0) (:152).  The formulas are complete on odd-order curves
// (P-256, P-384, BLS12-381 E(Fp)), so there are no exceptional inputs.
#pragma once
#include "mont.cuh"
#include "params_gen.cuh"

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
    static constexpr int SBITS = 256;
    ECB_DEV static u32 b(int i) { return P256_B[i]; }
    ECB_DEV static u32 b3(int i) { return P256_B3[i]; }
    ECB_DEV static u32 gx(int i) { return P256_GX[i]; }
    ECB_DEV static u32 gy(int i) { return P256_GY[i]; }
};
struct CurveP384 {
    typedef Mont<P384_FP> F;
    typedef Mont<P384_FN> FN;
    static constexpr bool A_M3 = true;
    static constexpr int FB = 48, SB = 48;
    static constexpr int SBITS = 384;
    ECB_DEV static u32 b(int i) { return P384_B[i]; }
    ECB_DEV static u32 b3(int i) { return P384_B3[i]; }
    ECB_DEV static u32 gx(int i) { return P384_GX[i]; }
    ECB_DEV static u32 gy(int i) { return P384_GY[i]; }
};
struct CurveBLSG1 {
    typedef Mont<BLS_FP> F;
    typedef Mont<BLS_FR> FN;
    static constexpr bool A_M3 = false;
    static constexpr int FB = 48, SB = 32;
    static constexpr int SBITS = 255;
    ECB_DEV static u32 b(int i) { return BLSG1_B[i]; }
    ECB_DEV static u32 b3(int i) { return BLSG1_B3[i]; }
    ECB_DEV static u32 gx(int i) { return BLSG1_GX[i]; }
    ECB_DEV static u32 gy(int i) { return BLSG1_GY[i]; }
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

    // complete addition
    ECB_DEV static void add(pt& r, const pt& p, const pt& q) {
        fe t0, t1, t2, t3, t4, X3, Y3, Z3, b;
        F::mul(t0, p.X, q.X);
        F::mul(t1, p.Y, q.Y);
        F::mul(t2, p.Z, q.Z);
        F::add(t3, p.X, p.Y);
        F::add(t4, q.X, q.Y);
        F::mul(t3, t3, t4);
        F::add(t4, t0, t1);
        F::sub(t3, t3, t4);
        F::add(t4, p.Y, p.Z);
        F::add(X3, q.Y, q.Z);
        F::mul(t4, t4, X3);
        F::add(X3, t1, t2);
        F::sub(t4, t4, X3);
        F::add(X3, p.X, p.Z);
        F::add(Y3, q.X, q.Z);
        F::mul(X3, X3, Y3);
        F::add(Y3, t0, t2);
        F::sub(Y3, X3, Y3);
        if (C::A_M3) {
            get_b(b);
            F::mul(Z3, b, t2);
            F::sub(X3, Y3, Z3);
            F::dbl(Z3, X3);
            F::add(X3, X3, Z3);
            F::sub(Z3, t1, X3);
            F::add(X3, t1, X3);
            F::mul(Y3, b, Y3);
            F::dbl(t1, t2);
            F::add(t2, t1, t2);
            F::sub(Y3, Y3, t2);
            F::sub(Y3, Y3, t0);
            F::dbl(t1, Y3);
            F::add(Y3, t1, Y3);
            F::dbl(t1, t0);
            F::add(t0, t1, t0);
            F::sub(t0, t0, t2);
            F::mul(t1, t4, Y3);
            F::mul(t2, t0, Y3);
            F::mul(Y3, X3, Z3);
            F::add(Y3, Y3, t2);
            F::mul(X3, t3, X3);
            F::sub(X3, X3, t1);
            F::mul(Z3, t4, Z3);
            F::mul(t1, t3, t0);
            F::add(Z3, Z3, t1);
        } else {
            get_b3(b);
            F::dbl(X3, t0);
            F::add(t0, X3, t0);
            F::mul(t2, b, t2);
            F::add(Z3, t1, t2);
            F::sub(t1, t1, t2);
            F::mul(Y3, b, Y3);
            F::mul(X3, t4, Y3);
            F::mul(t2, t3, t1);
            F::sub(X3, t2, X3);
            F::mul(Y3, Y3, t0);
            F::mul(t1, t1, Z3);
            F::add(Y3, t1, Y3);
            F::mul(t0, t0, t3);
            F::mul(Z3, Z3, t4);
            F::add(Z3, Z3, t0);
        }
        F::copy(r.X, X3);
        F::copy(r.Y, Y3);
        F::copy(r.Z, Z3);
    }

    // complete doubling
    ECB_DEV static void dbl(pt& r, const pt& p) {
        fe t0, t1, t2, t3, X3, Y3, Z3, b;
        if (C::A_M3) {
            get_b(b);
            F::sqr(t0, p.X);
            F::sqr(t1, p.Y);
            F::sqr(t2, p.Z);
            F::mul(t3, p.X, p.Y);
            F::dbl(t3, t3);
            F::mul(Z3, p.X, p.Z);
            F::dbl(Z3, Z3);
            F::mul(Y3, b, t2);
            F::sub(Y3, Y3, Z3);
            F::dbl(X3, Y3);
            F::add(Y3, X3, Y3);
            F::sub(X3, t1, Y3);
            F::add(Y3, t1, Y3);
            F::mul(Y3, X3, Y3);
            F::mul(X3, X3, t3);
            F::dbl(t3, t2);
            F::add(t2, t2, t3);
            F::mul(Z3, b, Z3);
            F::sub(Z3, Z3, t2);
            F::sub(Z3, Z3, t0);
            F::dbl(t3, Z3);
            F::add(Z3, Z3, t3);
            F::dbl(t3, t0);
            F::add(t0, t3, t0);
            F::sub(t0, t0, t2);
            F::mul(t0, t0, Z3);
            F::add(Y3, Y3, t0);
            F::mul(t0, p.Y, p.Z);
            F::dbl(t0, t0);
            F::mul(Z3, t0, Z3);
            F::sub(X3, X3, Z3);
            F::mul(Z3, t0, t1);
            F::dbl(Z3, Z3);
            F::dbl(Z3, Z3);
        } else {
            get_b3(b);
            F::sqr(t0, p.Y);
            F::dbl(Z3, t0);
            F::dbl(Z3, Z3);
            F::dbl(Z3, Z3);
            F::mul(t1, p.Y, p.Z);
            F::sqr(t2, p.Z);
            F::mul(t2, b, t2);
            F::mul(X3, t2, Z3);
            F::add(Y3, t0, t2);
            F::mul(Z3, t1, Z3);
            F::dbl(t1, t2);
            F::add(t2, t1, t2);
            F::sub(t0, t0, t2);
            F::mul(Y3, t0, Y3);
            F::add(Y3, X3, Y3);
            F::mul(t1, p.X, p.Y);
            F::mul(X3, t0, t1);
            F::dbl(X3, X3);
        }
        F::copy(r.X, X3);
        F::copy(r.Y, Y3);
        F::copy(r.Z, Z3);
    }
};

}  // namespace ecb
