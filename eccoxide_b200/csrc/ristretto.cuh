// ristretto.cuh — ristretto255 encodings around the edwards25519 kernels.
//
// The reference's RistrettoPoint (src/curve/curve25519/ristretto255.rs) wraps an edwards25519 Point:
// scale / mul_base are the Edwards scalar multiplications (:152-157), and the group lives in compress (:73) and
// decompress (:105), RFC 9496 4.3.1 / 4.3.2.  These are the batch forms of those two maps; the scalar
// multiplication in between is k_ed25519_mul / the comb.  Affine in, affine out: Z = 1, T = x y.
#pragma once
#include "kernels2.cuh"
#include "params_gen.cuh"

namespace ecb {

// RISTRETTO_INVSQRT_A_MINUS_D = 1 / sqrt(a - d), a = -1 (RFC 9496 Appendix A; ristretto255.rs:31) comes from params_gen.cuh:
// computed by tools/gen_params.py, pinned to the reference's bytes by the parity tests.

ECB_DEV u32 fe_is_negative(const fe25519& a) {   // low bit of the canonical representative (field_macros.rs:837)
    fe25519 f;
    F::freeze(f, a);
    return f.v[0] & 1u;
}
ECB_DEV void fe_abs(fe25519& r, const fe25519& a) {
    fe25519 n;
    F::neg(n, a);
    F::select(r, fe_is_negative(a), n, a);
}
// RFC 9496 4.2 SQRT_RATIO_M1 (FieldElement::sqrt_ratio_m1): r = |sqrt(u / v)| or |sqrt(i u / v)|; returns was_square
ECB_DEV u32 fe_sqrt_ratio_m1(fe25519& r, const fe25519& u, const fe25519& v) {
    fe25519 v3, v7, t, chk, nu, nui, sm1;
    F::from_words(sm1, ED25519_SQRTM1);
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
    F::mul(nui, nu, sm1);
    const u32 correct = F::eq(chk, u), flipped = F::eq(chk, nu), flipped_i = F::eq(chk, nui);
    F::mul(t, r, sm1);
    F::select(r, flipped | flipped_i, t, r);
    fe_abs(r, r);
    return correct | flipped;
}

// RistrettoPoint::decompress (ristretto255.rs:105-134): 32 bytes -> an edwards25519 representative, canonical
// affine x || y.  ok = 0 and x = y = 0 (not a curve point: a following ecb_ed25519_mul refuses it) otherwise.
ECB_DEV void ristretto255_decompress_body(size_t idx, const u32* enc, u32* out_xy, unsigned char* ok) {
    u32 w[8];
    ld_words<8>(w, enc + idx * 8);
    u32 good = F::is_canonical_words(w) & ((w[7] >> 31) ^ 1u);   // s < p (bit 255 clear is implied, checked anyway)
    fe25519 s, one, ss, u1, u2, u2s, v, d, inv, denx, deny, x, y, t;
    F::from_words(s, w);
    good &= (w[0] & 1u) ^ 1u;                                     // s non-negative
    F::set_one(one);
    F::from_words(d, ED25519_D);
    F::sqr(ss, s);
    F::sub(u1, one, ss);
    F::add(u2, one, ss);
    F::sqr(u2s, u2);
    F::sqr(t, u1);
    F::mul(t, t, d);
    F::neg(t, t);
    F::sub(v, t, u2s);
    F::mul(t, v, u2s);
    good &= fe_sqrt_ratio_m1(inv, one, t);
    F::mul(denx, inv, u2);
    F::mul(deny, inv, denx);
    F::mul(deny, deny, v);
    F::add(t, s, s);
    F::mul(x, t, denx);
    fe_abs(x, x);
    F::mul(y, u1, deny);
    F::mul(t, x, y);
    good &= fe_is_negative(t) ^ 1u;
    good &= F::is_zero(y) ^ 1u;
    F::freeze(x, x);
    F::freeze(y, y);
    const u32 m = good ? 0xffffffffu : 0u;
    ECB_UNROLL
    for (int i = 0; i < 8; i++) { x.v[i] &= m; y.v[i] &= m; }
    st_words<8>(out_xy + idx * 16, x.v);
    st_words<8>(out_xy + idx * 16 + 8, y.v);
    if (ok) ok[idx] = (unsigned char)good;
}

// RistrettoPoint::compress (ristretto255.rs:73-99) of the affine point (x, y): Z = 1, T = x y
ECB_DEV void ristretto255_compress_body(size_t idx, const u32* xy, u32* enc) {
    u32 w[16];
    ld_words<16>(w, xy + idx * 16);
    fe25519 x0, y0, one, t0, u1, u2, t, inv, den1, den2, zinv, ix, iy, ench, sm1, c, x, y, deninv, s;
    F::from_words(x0, w);
    F::from_words(y0, w + 8);
    F::set_one(one);
    F::from_words(sm1, ED25519_SQRTM1);
    F::from_words(c, RISTRETTO_INVSQRT_A_MINUS_D);
    F::mul(t0, x0, y0);
    F::add(u1, one, y0);
    F::sub(t, one, y0);
    F::mul(u1, u1, t);
    F::copy(u2, t0);                 // u2 = X Y with Z = 1
    F::sqr(t, u2);
    F::mul(t, t, u1);
    (void)fe_sqrt_ratio_m1(inv, one, t);
    F::mul(den1, inv, u1);
    F::mul(den2, inv, u2);
    F::mul(zinv, den1, den2);
    F::mul(zinv, zinv, t0);
    F::mul(ix, x0, sm1);
    F::mul(iy, y0, sm1);
    F::mul(ench, den1, c);
    F::mul(t, t0, zinv);
    const u32 rotate = fe_is_negative(t);
    F::select(x, rotate, iy, x0);
    F::select(y, rotate, ix, y0);
    F::select(deninv, rotate, ench, den2);
    F::mul(t, x, zinv);
    fe25519 ny;
    F::neg(ny, y);
    F::select(y, fe_is_negative(t), ny, y);
    F::sub(t, one, y);
    F::mul(s, deninv, t);
    fe_abs(s, s);
    F::freeze(s, s);
    st_words<8>(enc + idx * 8, s.v);
}

}  // namespace ecb
