// Host simulation of the device field layer (tests only; never linked into libeccbatch.so).
// Compiles the sm_100a headers with ECB_HOSTSIM (emulated carry flag) so the exact primitive
// sequences can be checked on a machine without a GPU.
#define ECB_HOSTSIM 1
#include "../../eccoxide_b200/csrc/fe25519.cuh"
#include "../../eccoxide_b200/csrc/mont_kinds.cuh"
#include "../../eccoxide_b200/csrc/fe43.cuh"
#include <string.h>
using namespace ecb;

template <class P> static void mont_op(int op, const u32* a, const u32* b, u32* r) {
    typedef Mont<P> F;
    typename F::el x, y, z;
    memcpy(x.v, a, 4 * P::N); memcpy(y.v, b, 4 * P::N);
    switch (op) {
        case 0: F::mul(z, x, y); break;
        case 1: F::sqr(z, x); break;
        case 2: F::add(z, x, y); break;
        case 3: F::sub(z, x, y); break;
        case 4: F::neg(z, x); break;
        case 5: F::to_mont(z, a); break;
        case 6: F::invert(z, x); break;
        case 7: F::from_mont(z.v, x); break;
        case 8: F::template mul_small<2>(z, x); break;
        case 9: F::template mul_small<3>(z, x); break;
        case 10: F::template mul_small<4>(z, x); break;
        case 11: F::template mul_small<8>(z, x); break;
        case 12: F::copy(z, x); F::template mul_small<8>(z, z); break;   // in place
        case 13: F::add_ct(z, x, y); break;
        case 14: F::sub_ct(z, x, y); break;
        case 15: F::neg_ct(z, x); break;
    }
    memcpy(r, z.v, 4 * P::N);
}
// merged three-operand forms: op 0 = a - b - c, 1 = a - b - 2c, 2 = the same with the result aliasing a
template <class P> static void mont_op3(int op, const u32* a, const u32* b, const u32* c, u32* r) {
    typedef Mont<P> F;
    typename F::el x, y, w, z;
    memcpy(x.v, a, 4 * P::N); memcpy(y.v, b, 4 * P::N); memcpy(w.v, c, 4 * P::N);
    switch (op) {
        case 0: F::sub2(z, x, y, w); break;
        case 1: F::sub_2x(z, x, y, w); break;
        case 2: F::sub2(x, x, y, w); F::copy(z, x); break;
        case 3: F::sub_2x(x, x, y, w); F::copy(z, x); break;
    }
    memcpy(r, z.v, 4 * P::N);
}

extern "C" {
// 30 divsteps on the low words: variant 0 = one step at a time (constant time), 1 = zero runs by ctz, 2 = jump table (modinv.cuh)
int hs_divsteps30(int variant, int zeta, u32 f0, u32 g0, int* t) {
    s32 u, v, q, r, z;
    if (variant == 0) z = sg_divsteps30(zeta, f0, g0, u, v, q, r);
    else if (variant == 1) z = sg_divsteps30_var(zeta, f0, g0, u, v, q, r);
    else z = sg_divsteps30_jump(zeta, f0, g0, u, v, q, r, SG_JUMP4);
    t[0] = u; t[1] = v; t[2] = q; t[3] = r;
    return z;
}
void hs_mul_full8(const u32* a, const u32* b, u32* t) { mul_full<8>(t, a, b); }
void hs_sqr_full8(const u32* a, u32* t) { sqr_full<8>(t, a); }
void hs_mul_full14(const u32* a, const u32* b, u32* t) { mul_full<14>(t, a, b); }
void hs_sqr_full14(const u32* a, u32* t) { sqr_full<14>(t, a); }

// op: 0 mul 1 sqr 2 add 3 sub 4 neg 5 freeze 6 invert 7 mul_small(b[0]) 8 pow_p58
void hs_fe25519(int op, const u32* a, const u32* b, u32* r) {
    fe25519 x, y, z;
    memcpy(x.v, a, 32); memcpy(y.v, b, 32);
    switch (op) {
        case 0: F25519::mul(z, x, y); break;
        case 1: F25519::sqr(z, x); break;
        case 2: F25519::add(z, x, y); break;
        case 3: F25519::sub(z, x, y); break;
        case 4: F25519::neg(z, x); break;
        case 5: F25519::freeze(z, x); break;
        case 6: F25519::invert(z, x); break;
        case 7: F25519::mul_small(z, x, b[0]); break;
        case 8: F25519::pow_p58(z, x); break;
        case 9: { fe25519 w; F25519::mul2(z, x, y, w, y, y); } break;    // first product of the interleaved pair
        case 10: { fe25519 w; F25519::mul2(w, x, y, z, y, y); } break;   // second product
    }
    memcpy(r, z.v, 32);
}
// FP64-pipe field (fe43.cuh): operands given as 8 words (converted with from_words), optionally loosened by
// adding `loosen` copies of the operand to itself (limbs up to (loosen + 1) * 2^43); result converted back with to_fe25519.
// op: 0 mul 1 sqr 2 add 3 sub 4 neg-then-mul (signed limbs) 5 conversion round trip
void hs_fe43(int op, const u32* a, const u32* b, int loosen, u32* r) {
    fe43 x, y, z;
    F43::from_words(x, a); F43::from_words(y, b);
    fe43 x0 = x, y0 = y;
    for (int i = 0; i < loosen; i++) { F43::add(x, x, x0); F43::add(y, y, y0); }
    switch (op) {
        case 0: F43::mul(z, x, y); break;
        case 1: F43::sqr(z, x); break;
        case 2: F43::add(z, x, y); break;
        case 3: F43::sub(z, x, y); break;
        case 4: { fe43 n; F43::neg(n, x); F43::sub(n, n, y0); F43::mul(z, n, y); } break;   // (-x - y0) * y
        case 5: z = x; break;
    }
    // chained use: feed the tight result through another product to exercise signed limbs
    fe25519 o;
    F43::to_fe25519(o, z);
    memcpy(r, o.v, 32);
}
double hs_fe43_maxlimb(int op, const u32* a, const u32* b, int loosen) {
    fe43 x, y, z;
    F43::from_words(x, a); F43::from_words(y, b);
    fe43 x0 = x, y0 = y;
    for (int i = 0; i < loosen; i++) { F43::add(x, x, x0); F43::add(y, y, y0); }
    if (op == 0) F43::mul(z, x, y); else F43::sqr(z, x);
    double m = 0;
    for (int i = 0; i < 6; i++) { double t = z.v[i] < 0 ? -z.v[i] : z.v[i]; if (t > m) m = t; }
    return m;
}
u32 hs_fe25519_is_canonical(const u32* a) { return F25519::is_canonical_words(a); }

// field: 0 P256_FP 1 P256_FN 2 P384_FP 3 P384_FN 4 BLS_FP 5 BLS_FR 6 K256_FP (plain, pseudo-Mersenne) 7 K256_FN
void hs_mont(int field, int op, const u32* a, const u32* b, u32* r) {
    switch (field) {
        case 0: mont_op<P256_FP>(op, a, b, r); break;
        case 1: mont_op<P256_FN>(op, a, b, r); break;
        case 2: mont_op<P384_FP>(op, a, b, r); break;
        case 3: mont_op<P384_FN>(op, a, b, r); break;
        case 4: mont_op<BLS_FP>(op, a, b, r); break;
        case 5: mont_op<BLS_FR>(op, a, b, r); break;
        case 6: mont_op<K256_FP>(op, a, b, r); break;
        case 7: mont_op<K256_FN>(op, a, b, r); break;
    }
}
void hs_mont3(int field, int op, const u32* a, const u32* b, const u32* c, u32* r) {
    switch (field) {
        case 0: mont_op3<P256_FP>(op, a, b, c, r); break;
        case 1: mont_op3<P256_FN>(op, a, b, c, r); break;
        case 2: mont_op3<P384_FP>(op, a, b, c, r); break;
        case 3: mont_op3<P384_FN>(op, a, b, c, r); break;
        case 4: mont_op3<BLS_FP>(op, a, b, c, r); break;
        case 5: mont_op3<BLS_FR>(op, a, b, c, r); break;
        case 6: mont_op3<K256_FP>(op, a, b, c, r); break;
        case 7: mont_op3<K256_FN>(op, a, b, c, r); break;
    }
}
}
