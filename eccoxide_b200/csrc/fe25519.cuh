// fe25519.cuh — GF(2^255-19) on 8 saturated 32-bit limbs.
//
// Replaces (for the batch path) the reference's 5x51-bit fiat backend
// src/curve/fiat/curve25519_64.rs:217 (carry_mul), :285 (carry_square), :379 (add),
// :400 (sub), :421 (opp), :482 (to_bytes, canonicalising), :631 (from_bytes, which
// accepts values >= p) and the wrappers src/curve/curve25519.rs:62-117, plus the
// exponentiation chains src/curve/curve25519.rs:155-195 (pow_2_250_m1,
// invert_or_zero, pow_p58).
//
// Representation: any 256-bit value congruent to the element ("loose"); 2^256 = 38
// (mod p) is folded wherever a carry/borrow leaves the top limb.  Only freeze()
// produces the canonical representative in [0, p), which is what the reference's
// to_bytes emits.  mul = 64 + 8 IMAD.WIDE.U32, sqr = 36 + 8.
#pragma once
#include "limb.cuh"
#include "modinv.cuh"

namespace ecb {

struct fe25519 {
    u32 v[8];
};

struct F25519 {
    typedef fe25519 el;
    static constexpr int N = 8;
    static constexpr bool BLOCK_INV = true;   // k_batch_inv (one inversion per block) measured faster for this field

    ECB_DEV static void set_zero(el& r) {
        ECB_UNROLL
        for (int i = 0; i < 8; i++) r.v[i] = 0;
    }
    ECB_DEV static void set_u32(el& r, u32 x) {
        set_zero(r);
        r.v[0] = x;
    }
    ECB_DEV static void set_one(el& r) { set_u32(r, 1); }
    ECB_DEV static void copy(el& r, const el& a) {
        ECB_UNROLL
        for (int i = 0; i < 8; i++) r.v[i] = a.v[i];
    }

    // fold a 512-bit product into 8 limbs: lo + 38*hi, twice
    ECB_DEV static void fold(el& r, const u32* t) {
        u32 R[10];
        ECB_UNROLL
        for (int i = 0; i < 8; i++) R[i] = t[i];
        R[8] = 0;
        R[9] = 0;
        mac_chain<4, true>(R, t + 8, 38u);       // 38*t[8,10,12,14] at limbs 0,2,4,6
        mac_chain<4, false>(R + 1, t + 9, 38u);  // 38*t[9,11,13,15] at limbs 1,3,5,7
        // R[8] <= 38 : second fold
        R[0] = mad_lo_cc(R[8], 38u, R[0]);
        ECB_UNROLL
        for (int i = 1; i < 8; i++) R[i] = addc_cc(R[i], 0);
        u32 c = addc(0, 0);
        R[0] += 38u * c;  // a wrapped value is tiny, this cannot carry
        ECB_UNROLL
        for (int i = 0; i < 8; i++) r.v[i] = R[i];
    }

    ECB_DEV static void mul(el& r, const el& a, const el& b) {
        u32 t[16];
        mul_full<8>(t, a.v, b.v);
        fold(r, t);
    }
    ECB_DEV static void sqr(el& r, const el& a) {
        u32 t[16];
        sqr_full<8>(t, a.v);
        fold(r, t);
    }
    // r1 = a1 * b1 and r2 = a2 * b2, rows interleaved (limb.cuh mul_full2): for latency-bound callers
    ECB_DEV static void mul2(el& r1, const el& a1, const el& b1, el& r2, const el& a2, const el& b2) {
        u32 t1[16], t2[16];
        mul_full2<8>(t1, a1.v, b1.v, t2, a2.v, b2.v);
        fold(r1, t1);
        fold(r2, t2);
    }
    // out-of-line copies (operands by value, so they stay in registers) for the kernels whose window
    // loop would not fit the instruction cache with every product inlined
    ECB_DEVNI static el mul_v(el a, el b) {
        el r;
        mul(r, a, b);
        return r;
    }
    ECB_DEVNI static el sqr_v(el a) {
        el r;
        sqr(r, a);
        return r;
    }
    ECB_DEV static void mul_ni(el& r, const el& a, const el& b) { r = mul_v(a, b); }
    ECB_DEV static void sqr_ni(el& r, const el& a) { r = sqr_v(a); }
    // r = a * k for a small constant k (< 2^26)
    ECB_DEV static void mul_small(el& r, const el& a, u32 k) {
        u32 R[10];
        ECB_UNROLL
        for (int i = 0; i < 10; i++) R[i] = 0;
        mac_chain<4, true>(R, a.v, k);
        mac_chain<4, false>(R + 1, a.v + 1, k);
        R[0] = mad_lo_cc(R[8], 38u, R[0]);
        ECB_UNROLL
        for (int i = 1; i < 8; i++) R[i] = addc_cc(R[i], 0);
        u32 c = addc(0, 0);
        R[0] += 38u * c;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) r.v[i] = R[i];
    }

    ECB_DEV static void add(el& r, const el& a, const el& b) {
        u32 t[8];
        u32 c = add_n<8>(t, a.v, b.v);
        t[0] = add_cc(t[0], 38u * c);
        ECB_UNROLL
        for (int i = 1; i < 8; i++) t[i] = addc_cc(t[i], 0);
        u32 c2 = addc(0, 0);
        t[0] += 38u * c2;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) r.v[i] = t[i];
    }
    ECB_DEV static void sub(el& r, const el& a, const el& b) {
        u32 t[8];
        u32 c = sub_n<8>(t, a.v, b.v);
        t[0] = sub_cc(t[0], 38u * c);
        ECB_UNROLL
        for (int i = 1; i < 8; i++) t[i] = subc_cc(t[i], 0);
        u32 c2 = subc(0, 0) & 1;
        t[0] -= 38u * c2;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) r.v[i] = t[i];
    }
    ECB_DEV static void neg(el& r, const el& a) {
        el z;
        set_zero(z);
        sub(r, z, a);
    }
    ECB_DEV static void dbl(el& r, const el& a) { add(r, a, a); }

    // canonical representative in [0, p)
    ECB_DEV static void freeze(el& r, const el& a) {
        u32 t[8];
        u32 top = a.v[7] >> 31;
        t[0] = add_cc(a.v[0], 19u * top);
        ECB_UNROLL
        for (int i = 1; i < 7; i++) t[i] = addc_cc(a.v[i], 0);
        t[7] = addc(a.v[7] & 0x7fffffffu, 0);
        // t in [0, 2^255 + 19); subtract p iff t >= p  <=>  bit 255 of t+19 set
        u32 s[8];
        s[0] = add_cc(t[0], 19u);
        ECB_UNROLL
        for (int i = 1; i < 8; i++) s[i] = addc_cc(t[i], 0);
        u32 ge = s[7] >> 31;
        s[7] &= 0x7fffffffu;
        u32 m = 0u - ge;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) r.v[i] = (s[i] & m) | (t[i] & ~m);
    }
    ECB_DEV static u32 is_zero(const el& a) {  // 1 if a == 0 (mod p)
        el f;
        freeze(f, a);
        u32 o = 0;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) o |= f.v[i];
        return o == 0 ? 1u : 0u;
    }
    ECB_DEV static u32 eq(const el& a, const el& b) {
        el d;
        sub(d, a, b);
        return is_zero(d);
    }
    // r = c ? a : b   (c in {0,1})
    ECB_DEV static void select(el& r, u32 c, const el& a, const el& b) {
        u32 m = 0u - c;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) r.v[i] = (a.v[i] & m) | (b.v[i] & ~m);
    }
    ECB_DEV static void cswap(u32 c, el& a, el& b) {
        u32 m = 0u - c;
        ECB_UNROLL
        for (int i = 0; i < 8; i++) {
            u32 x = (a.v[i] ^ b.v[i]) & m;
            a.v[i] ^= x;
            b.v[i] ^= x;
        }
    }

    ECB_DEV static void sqr_n(el& r, const el& a, int n) {
        sqr(r, a);
        for (int i = 1; i < n; i++) sqr(r, r);
    }
    // (a^(2^250-1), a^11): the shared ref10 prefix, curve25519.rs:155
    ECB_DEV static void pow_2_250_m1(el& t250, el& z11, const el& z1) {
        el z2, z9, t, z_5_0, z_10_0, z_20_0, z_50_0, z_100_0;
        sqr(z2, z1);
        sqr_n(t, z2, 2);        // z8
        mul(z9, t, z1);
        mul(z11, z9, z2);
        sqr(t, z11);            // z22
        mul(z_5_0, t, z9);
        sqr_n(t, z_5_0, 5);
        mul(z_10_0, t, z_5_0);
        sqr_n(t, z_10_0, 10);
        mul(z_20_0, t, z_10_0);
        sqr_n(t, z_20_0, 20);
        mul(t, t, z_20_0);
        sqr_n(t, t, 10);
        mul(z_50_0, t, z_10_0);
        sqr_n(t, z_50_0, 50);
        mul(z_100_0, t, z_50_0);
        sqr_n(t, z_100_0, 100);
        mul(t, t, z_100_0);
        sqr_n(t, t, 50);
        mul(t250, t, z_50_0);
    }
    // a^(p-2); 0 -> 0 (curve25519.rs:191 invert_or_zero): the reference's Fermat chain
    ECB_DEV static void invert_fermat(el& r, const el& a) {
        el t, z11;
        pow_2_250_m1(t, z11, a);
        sqr_n(t, t, 5);
        mul(r, t, z11);
    }
    // same value by safegcd divsteps (modinv.cuh): what the batch-inversion kernels use; 0 -> 0
    ECB_DEV static void invert(el& r, const el& a) {
        el c;
        freeze(c, a);
        const u32 p[8] = {0xffffffedu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x7fffffffu};
        sg_modinv<8, 9, 22>(r.v, c.v, p);
    }
#ifndef ECB_HOSTSIM
    // the same inverse by one whole warp (every lane passes the same a): ~3x shorter latency, for the
    // block-level batch inversions where one inversion runs while the rest of the block waits
    __device__ __forceinline__ static void invert_warp(el& r, const el& a, const u32* jump) {
        el c;
        freeze(c, a);
        const u32 p[8] = {0xffffffedu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x7fffffffu};
        sg_modinv_warp<8, 9, 22>(r.v, c.v, p, jump);
    }
#endif
    // a^((p-5)/8) (curve25519.rs:185)
    ECB_DEV static void pow_p58(el& r, const el& a) {
        el t, z11;
        pow_2_250_m1(t, z11, a);
        sqr_n(t, t, 2);
        mul(r, t, a);
    }

    // wire format: 32 bytes little-endian = the 8 limbs in order.  No canonical
    // check here (from_bytes_unchecked_le semantics); is_canonical() is separate.
    ECB_DEV static void from_words(el& r, const u32* w) {
        ECB_UNROLL
        for (int i = 0; i < 8; i++) r.v[i] = w[i];
    }
    ECB_DEV static u32 is_canonical_words(const u32* w) {  // value < p ?
        // w >= p  <=>  w + 19 >= 2^255 (for w < 2^256, also true when bit 255 set)
        u32 s = add_cc(w[0], 19u);
        ECB_UNROLL
        for (int i = 1; i < 8; i++) s = addc_cc(w[i], 0);
        u32 c = addc(0, 0);
        return ((s >> 31) | c) ? 0u : 1u;
    }
};

// curve constants (src/curve/curve25519.rs:364-423), little-endian 32-bit limbs
ECB_CONST u32 ED25519_D[8] = {0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du,
                              0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu};
ECB_CONST u32 ED25519_D2[8] = {0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au,
                               0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu};
ECB_CONST u32 ED25519_SQRTM1[8] = {0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u,
                                   0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u};
ECB_CONST u32 ED25519_BX[8] = {0x8f25d51au, 0xc9562d60u, 0x9525a7b2u, 0x692cc760u,
                               0xfdd6dc5cu, 0xc0a4e231u, 0xcd6e53feu, 0x216936d3u};
ECB_CONST u32 ED25519_BY[8] = {0x66666658u, 0x66666666u, 0x66666666u, 0x66666666u,
                               0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u};

}  // namespace ecb
