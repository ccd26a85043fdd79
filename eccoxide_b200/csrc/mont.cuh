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
#include "modinv.cuh"

namespace ecb {

template <int N_>
struct fe_mont {
    u32 v[N_];
};

// Per-prime reduction strategy.  kind 0 = generic CIOS (2 N^2 IMAD.WIDE per product).
// kind 1 = p256: -p^-1 = 1 (mod 2^96) and p = 2^256 - 2^224 + 2^192 + 2^96 - 1, so the Montgomery
// reduction of a 512-bit product is three rounds of shifted additions / subtractions of
// m = T mod 2^96 — no multiplication at all; product = N^2, square = N(N+1)/2 IMAD.WIDE.
// kind 2 = p384: -p^-1 = 1 + 2^32 + 2^64 (mod 2^96) and p = 2^384 - 2^128 - 2^96 + 2^32 - 1: four
// rounds of 96 bits, each a handful of shifted additions / subtractions of m — again no multiply.
// kind 3 = secp256k1's p = 2^256 - 2^32 - 977: NO Montgomery domain (the parameter class gives R = 1, so "to_mont" and
// "r2" are identities): a 512-bit product folds as lo + hi (2^32 + 977) — eight IMAD.WIDE and a shifted addition —
// product = N^2 + N + 1, square = N(N+1)/2 + N + 1 instead of the 2 N^2 / N(N+1)/2 + N^2 of the generic rows.
template <class P>
struct MontKind {
    static constexpr int kind = 0;
};

// P supplies: N, INV (= -p^-1 mod 2^32) and accessors mod(i), r1(i) (= R mod p, "one"),
// r2(i) (= R^2 mod p) over little-endian 32-bit limbs held in `__constant__ const`
// arrays; with the loops unrolled the limbs fold to immediates.
template <class P>
struct Mont {
    static constexpr int N = P::N;
    static constexpr bool BLOCK_INV = false;  // wider fields: the per-thread form of the batch inversion is faster (registers)
    typedef P P_;
    typedef fe_mont<P::N> el;
    // p256r1 and p384r1 keep elements "loose": any representative below 2^(32 N), not necessarily
    // below p (see the note above add()); the generic primes stay canonical
    static constexpr bool LOOSE = MontKind<P>::kind == 1 || MontKind<P>::kind == 2 || MontKind<P>::kind == 3;

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

    // ---- p256: one round of T += m * p * 2^(32 O), m = T[O .. O+ML) (ML = 3, or 2 in the last
    // round); limbs O .. O+ML-1 become zero and are simply skipped.
    //   m*p = m*2^256 - m*2^224 + m*2^192 + m*2^96 - m
    template <int O, int ML>
    ECB_DEV static void p256_round(u32* T) {
        const u32 m0 = T[O], m1 = T[O + 1], m2 = (ML == 3) ? T[O + 2] : 0u;
        // U = m * (2^96 + 2^192 + 2^256), relative limbs 3..11 (3..7 are plain copies of m)
        u32 U8 = add_cc(m2, m0);
        u32 U9 = addc_cc(m1, 0u);
        u32 U10 = addc_cc(m2, 0u);
        u32 U11 = addc(0u, 0u);
        // V = U - m * 2^224 (relative limbs 7..11); V >= 0
        u32 V7 = sub_cc(m1, m0);
        u32 V8 = subc_cc(U8, m1);
        u32 V9 = subc_cc(U9, m2);
        u32 V10 = subc_cc(U10, 0u);
        u32 V11 = subc(U11, 0u);
        T[O + 3] = add_cc(T[O + 3], m0);
        T[O + 4] = addc_cc(T[O + 4], m1);
        T[O + 5] = addc_cc(T[O + 5], m2);
        T[O + 6] = addc_cc(T[O + 6], m0);
        T[O + 7] = addc_cc(T[O + 7], V7);
        T[O + 8] = addc_cc(T[O + 8], V8);
        T[O + 9] = addc_cc(T[O + 9], V9);
        T[O + 10] = addc_cc(T[O + 10], V10);
        if (O + 11 <= 16) T[O + 11] = addc_cc(T[O + 11], V11);
        ECB_UNROLL
        for (int j = O + 12; j <= 16; j++) T[j] = addc_cc(T[j], 0u);
        (void)addc(0u, 0u);
    }
    // T: 17 limbs, T[16] = 0 on entry, value < p^2.  r = T / 2^256 mod p, canonical.
    ECB_DEV static void p256_reduce(el& r, u32* T) {
        p256_round<0, 3>(T);
        p256_round<3, 3>(T);
        p256_round<6, 2>(T);
        // T[8..16] < 2^256 + p: fold the carry word by adding 2^256 - p once; the result is < 2^256
        // ("loose": congruent, not necessarily < p — see the representation note above add())
        fold_carry(r, T + 8, T[16]);
    }
    // r = t + c * (2^256 - p) mod 2^256 for c in {0, 1};  2^256 - p = 2^224 - 2^192 - 2^96 + 1
    // acc (optional): the lazy forms below pass a counter; the carry is then ADDED to it instead of being returned alone
    ECB_DEV static u32 fold_carry(el& r, const u32* t, u32 c, const u32* acc = nullptr) {
        const u32 m = 0u - c;
        if constexpr (MontKind<P>::kind == 3) {
            // 2^256 - p = 2^32 + 977 = limbs {977, 1, 0, ...}; valid for any small c (977 c < 2^32)
            r.v[0] = add_cc(t[0], 977u * c);
            r.v[1] = addc_cc(t[1], c);
            ECB_UNROLL
            for (int i = 2; i < N; i++) r.v[i] = addc_cc(t[i], 0u);
            return acc ? addc(*acc, 0u) : addc(0u, 0u);
        }
        if constexpr (MontKind<P>::kind == 2) {
            // 2^384 - p = 2^128 + 2^96 - 2^32 + 1 = limbs {1, ffffffff, ffffffff, 0, 1, 0, ...}
            r.v[0] = add_cc(t[0], c);
            r.v[1] = addc_cc(t[1], m);
            r.v[2] = addc_cc(t[2], m);
            r.v[3] = addc_cc(t[3], 0u);
            r.v[4] = addc_cc(t[4], c);
            ECB_UNROLL
            for (int i = 5; i < N; i++) r.v[i] = addc_cc(t[i], 0u);
            return acc ? addc(*acc, 0u) : addc(0u, 0u);
        }
        r.v[0] = add_cc(t[0], c);
        r.v[1] = addc_cc(t[1], 0u);
        r.v[2] = addc_cc(t[2], 0u);
        r.v[3] = addc_cc(t[3], m);
        r.v[4] = addc_cc(t[4], m);
        r.v[5] = addc_cc(t[5], m);
        r.v[6] = addc_cc(t[6], m << 1);
        r.v[7] = addc_cc(t[7], 0u);
        return acc ? addc(*acc, 0u) : addc(0u, 0u);
    }
    // r = t - c * (2^256 - p) mod 2^256 ; returns the borrow
    ECB_DEV static u32 fold_borrow(el& r, const u32* t, u32 c, const u32* acc = nullptr) {
        const u32 m = 0u - c;
        if constexpr (MontKind<P>::kind == 3) {
            r.v[0] = sub_cc(t[0], 977u * c);
            r.v[1] = subc_cc(t[1], c);
            ECB_UNROLL
            for (int i = 2; i < N; i++) r.v[i] = subc_cc(t[i], 0u);
            return acc ? subc(*acc, 0u) : 0u - subc(0u, 0u);
        }
        if constexpr (MontKind<P>::kind == 2) {
            r.v[0] = sub_cc(t[0], c);
            r.v[1] = subc_cc(t[1], m);
            r.v[2] = subc_cc(t[2], m);
            r.v[3] = subc_cc(t[3], 0u);
            r.v[4] = subc_cc(t[4], c);
            ECB_UNROLL
            for (int i = 5; i < N; i++) r.v[i] = subc_cc(t[i], 0u);
            return acc ? subc(*acc, 0u) : 0u - subc(0u, 0u);   // 0 or 1, written so that its range stays unknown to the compiler (refold_borrow must remain a loop); lazy: the counter runs DOWN
        }
        r.v[0] = sub_cc(t[0], c);
        r.v[1] = subc_cc(t[1], 0u);
        r.v[2] = subc_cc(t[2], 0u);
        r.v[3] = subc_cc(t[3], m);
        r.v[4] = subc_cc(t[4], m);
        r.v[5] = subc_cc(t[5], m);
        r.v[6] = subc_cc(t[6], m << 1);
        r.v[7] = subc_cc(t[7], 0u);
        return acc ? subc(*acc, 0u) : 0u - subc(0u, 0u);   // 0 or 1, written so that its range stays unknown to the compiler (refold_borrow must remain a loop); lazy: the counter runs DOWN
    }

    // ---- loose representation, small multiples.  r = t + c (2^(32N) - p) mod 2^(32N) for a small c (0 <= c < 2^16),
    // returns the carry out.  The limbs of c (2^256 - p) = c 2^224 - c 2^192 - c 2^96 + c are
    // {c, 0, 0, -c, ~0, ~0, ~c, c - 1} for c >= 1 (all zero for c = 0); of c (2^384 - p) = c 2^128 + c 2^96 - c 2^32 + c:
    // {c, -c, ~0, c - 1, c, 0, ...}.
    ECB_DEV static u32 fold_carry_small(el& r, const u32* t, u32 c, const u32* acc = nullptr) {
        if constexpr (MontKind<P>::kind == 3) return fold_carry(r, t, c, acc);
        const u32 nz = 0u - (u32)(c != 0u);
        if constexpr (MontKind<P>::kind == 2) {
            r.v[0] = add_cc(t[0], c);
            r.v[1] = addc_cc(t[1], 0u - c);
            r.v[2] = addc_cc(t[2], nz);
            r.v[3] = addc_cc(t[3], (c - 1u) & nz);
            r.v[4] = addc_cc(t[4], c);
            ECB_UNROLL
            for (int i = 5; i < N; i++) r.v[i] = addc_cc(t[i], 0u);
            return acc ? addc(*acc, 0u) : addc(0u, 0u);
        }
        r.v[0] = add_cc(t[0], c);
        r.v[1] = addc_cc(t[1], 0u);
        r.v[2] = addc_cc(t[2], 0u);
        r.v[3] = addc_cc(t[3], 0u - c);
        r.v[4] = addc_cc(t[4], nz);
        r.v[5] = addc_cc(t[5], nz);
        r.v[6] = addc_cc(t[6], ~c & nz);
        r.v[7] = addc_cc(t[7], (c - 1u) & nz);
        return acc ? addc(*acc, 0u) : addc(0u, 0u);
    }
    // r = t - c (2^(32N) - p) mod 2^(32N), returns the borrow
    ECB_DEV static u32 fold_borrow_small(el& r, const u32* t, u32 c, const u32* acc = nullptr) {
        if constexpr (MontKind<P>::kind == 3) return fold_borrow(r, t, c, acc);
        const u32 nz = 0u - (u32)(c != 0u);
        if constexpr (MontKind<P>::kind == 2) {
            r.v[0] = sub_cc(t[0], c);
            r.v[1] = subc_cc(t[1], 0u - c);
            r.v[2] = subc_cc(t[2], nz);
            r.v[3] = subc_cc(t[3], (c - 1u) & nz);
            r.v[4] = subc_cc(t[4], c);
            ECB_UNROLL
            for (int i = 5; i < N; i++) r.v[i] = subc_cc(t[i], 0u);
            return acc ? subc(*acc, 0u) : 0u - subc(0u, 0u);   // 0 or 1, written so that its range stays unknown to the compiler (refold_borrow must remain a loop); lazy: the counter runs DOWN
        }
        r.v[0] = sub_cc(t[0], c);
        r.v[1] = subc_cc(t[1], 0u);
        r.v[2] = subc_cc(t[2], 0u);
        r.v[3] = subc_cc(t[3], 0u - c);
        r.v[4] = subc_cc(t[4], nz);
        r.v[5] = subc_cc(t[5], nz);
        r.v[6] = subc_cc(t[6], ~c & nz);
        r.v[7] = subc_cc(t[7], (c - 1u) & nz);
        return acc ? subc(*acc, 0u) : 0u - subc(0u, 0u);   // 0 or 1, written so that its range stays unknown to the compiler (refold_borrow must remain a loop); lazy: the counter runs DOWN
    }
    // The fold itself can carry once more (t within c 2^224 of 2^(32N)): about one operand pair in 2^32.  The repeat
    // is a LOOP so that the assembler keeps it a branch: written as `if`, it becomes eight predicated instructions
    // that every field addition then issues for nothing (6 % of the issue slots of a P-256 doubling).  The second fold
    // cannot carry (the value is below c 2^224 by then).
    ECB_DEV static void refold_carry(el& r, u32 c2) {
        ECB_NOUNROLL
        for (; c2 != 0u; c2--) (void)fold_carry(r, r.v, 1u);
    }
    ECB_DEV static void refold_borrow(el& r, u32 b2) {
        ECB_NOUNROLL
        for (; b2 != 0u; b2--) (void)fold_borrow(r, r.v, 1u);
    }

    // ---- p384: one round at limb offset O.  t = T[O..O+3); m = t (1 + 2^32 + 2^64) mod 2^96;
    // T += m p 2^(32 O) with m p = m 2^384 - m 2^128 - m 2^96 + m (2^32 - 1).
    // t + (m (2^32 - 1) mod 2^96) = k 2^96 with k = [t != 0], so limbs O..O+2 vanish and the rest is
    //   V = (hi32(m (2^32 - 1)) + k) + m 2^(32*9) - m (1 + 2^32)        (relative limbs 3..14, V >= 0;
    //   m (1 + 2^32) has five limbs)
    template <int O>
    ECB_DEV static void p384_round(u32* T) {
        const u32 t0 = T[O], t1 = T[O + 1], t2 = T[O + 2];
        const u32 m0 = t0;
        const u32 m1 = add_cc(t1, t0);
        const u32 a2 = addc(t2, t1);
        const u32 m2 = a2 + t0;
        const u32 k = (t0 | t1 | t2) ? 1u : 0u;
        // d3 = top limb of m (2^32 - 1) = [0, m0, m1, m2] - [m0, m1, m2, 0]
        (void)sub_cc(0u, m0);
        (void)subc_cc(m0, m1);
        (void)subc_cc(m1, m2);
        const u32 d3 = subc(m2, 0u);
        const u32 e3 = add_cc(d3, k);
        const u32 ce = addc(0u, 0u);
        // Neg = m (1 + 2^32): 4 limbs
        const u32 n1 = add_cc(m1, m0);
        const u32 n2 = addc_cc(m2, m1);
        const u32 n3 = addc_cc(m2, 0u);
        const u32 n4 = addc(0u, 0u);
        // V = [e3, ce, 0, 0, 0.., m0, m1, m2 at rel 12..14] - [m0, n1, n2, n3, n4]
        const u32 v3 = sub_cc(e3, m0);
        const u32 v4 = subc_cc(ce, n1);
        const u32 v5 = subc_cc(0u, n2);
        const u32 v6 = subc_cc(0u, n3);
        const u32 v7 = subc_cc(0u, n4);
        const u32 vm = subc_cc(0u, 0u);   // rel 8..11: 0 - borrow, the borrow rides through unchanged
        const u32 v12 = subc_cc(m0, 0u);
        const u32 v13 = subc_cc(m1, 0u);
        const u32 v14 = subc(m2, 0u);
        T[O + 3] = add_cc(T[O + 3], v3);
        T[O + 4] = addc_cc(T[O + 4], v4);
        T[O + 5] = addc_cc(T[O + 5], v5);
        T[O + 6] = addc_cc(T[O + 6], v6);
        T[O + 7] = addc_cc(T[O + 7], v7);
        ECB_UNROLL
        for (int j = 8; j <= 11; j++) T[O + j] = addc_cc(T[O + j], vm);
        T[O + 12] = addc_cc(T[O + 12], v12);
        T[O + 13] = addc_cc(T[O + 13], v13);
        T[O + 14] = addc_cc(T[O + 14], v14);
        ECB_UNROLL
        for (int j = O + 15; j <= 24; j++) T[j] = addc_cc(T[j], 0u);
        (void)addc(0u, 0u);
    }
    // T: 25 limbs, T[24] = 0 on entry, value < 2^768.  r = T / 2^384 mod p, loose (< 2^384).
    ECB_DEV static void p384_reduce(el& r, u32* T) {
        p384_round<0>(T);
        p384_round<3>(T);
        p384_round<6>(T);
        p384_round<9>(T);
        // T[12..24] < 2^384 + p: fold the carry word by adding 2^384 - p once; the result is < 2^384 (loose)
        fold_carry(r, T + 12, T[24]);
    }

    // ---- secp256k1: T (16 limbs) -> lo + hi (2^32 + 977), twice, loose result (< 2^256)
    ECB_DEV static void k256_reduce(el& r, const u32* T) {
        u32 R[10];
        ECB_UNROLL
        for (int i = 0; i < 8; i++) R[i] = T[i];
        R[8] = 0;
        R[9] = 0;
        mac_chain<4, true>(R, T + 8, 977u);        // 977 * T[8, 10, 12, 14] at limbs 0, 2, 4, 6 (carry into R[8])
        mac_chain<4, false>(R + 1, T + 9, 977u);   // 977 * T[9, 11, 13, 15] at limbs 1, 3, 5, 7: R[8] < 2^11, no carry out
        R[1] = add_cc(R[1], T[8]);                 // + hi << 32
        ECB_UNROLL
        for (int i = 2; i <= 8; i++) R[i] = addc_cc(R[i], T[i + 7]);
        R[9] = addc(0u, 0u);
        // V = R[8] + 2^32 R[9] < 2^33:  V (2^32 + 977) = (977 R[8]) + 2^32 (977 R[9] + R[8]) + 2^64 R[9]
        const u32 w0 = mul_lo(R[8], 977u);
        const u64 a1 = (u64)mul_hi(R[8], 977u) + (u64)(977u * R[9]) + R[8];
        u32 t[8];
        t[0] = add_cc(R[0], w0);
        t[1] = addc_cc(R[1], (u32)a1);
        t[2] = addc_cc(R[2], (u32)(a1 >> 32) + R[9]);
        ECB_UNROLL
        for (int i = 3; i < 8; i++) t[i] = addc_cc(R[i], 0u);
        const u32 c = addc(0u, 0u);
        // a wrapped value is below 2^66: adding 2^32 + 977 once more cannot carry out of limb 2
        r.v[0] = add_cc(t[0], 977u * c);
        r.v[1] = addc_cc(t[1], c);
        r.v[2] = addc(t[2], 0u);
        ECB_UNROLL
        for (int i = 3; i < 8; i++) r.v[i] = t[i];
    }

    ECB_DEV static void mul(el& r, const el& a, const el& b) {
        if constexpr (MontKind<P>::kind == 3) {
            u32 T[16];
            mul_full<N>(T, a.v, b.v);
            k256_reduce(r, T);
            return;
        }
        if constexpr (MontKind<P>::kind == 2) {
            u32 T[25];
            mul_full<N>(T, a.v, b.v);
            T[24] = 0;
            p384_reduce(r, T);
            return;
        }
        if constexpr (MontKind<P>::kind == 1) {
            u32 T[17];
            mul_full<N>(T, a.v, b.v);
            T[16] = 0;
            p256_reduce(r, T);
            return;
        }
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
    // Montgomery reduction of a full 2N-limb product T (generic prime): the CIOS reduction rows of
    // mul() on their own — per row m = w0 * INV, window += m * p, shift one limb, pull in the next
    // limb of T at the top.  N^2 IMAD.WIDE; with sqr_full this makes a square N(N+1)/2 + N^2 instead
    // of the 2 N^2 of mul(a, a).
    ECB_DEV static void redc_full(el& r, const u32* T) {
        u32 ev[N + 2], od[N + 2];
        ECB_UNROLL
        for (int i = 0; i < N; i++) { ev[i] = T[i]; od[i] = 0; }
        ev[N] = ev[N + 1] = od[N] = od[N + 1] = 0;
        u32 L = 0;
        u32 pm[N];
        ECB_UNROLL
        for (int i = 0; i < N; i++) pm[i] = P::mod(i);
        ECB_UNROLL
        for (int i = 0; i < N; i++) {
            ev[0] = add_cc(ev[0], L);            // its carry enters the odd chain (limb 1)
            u32 m = ev[0] * P::INV;
            od[0] = madc_lo_cc(pm[1], m, od[0]);
            od[1] = madc_hi_cc(pm[1], m, od[1]);
            ECB_UNROLL
            for (int k = 1; k < N / 2; k++) {
                od[2 * k] = madc_lo_cc(pm[2 * k + 1], m, od[2 * k]);
                od[2 * k + 1] = madc_hi_cc(pm[2 * k + 1], m, od[2 * k + 1]);
            }
            od[N] = addc(od[N], 0);
            mac_chain<N / 2, true>(ev, pm, m);
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
            ev[N - 1] = add_cc(ev[N - 1], T[N + i]);   // next limb of T enters at the top of the window
            ev[N] = addc(ev[N], 0);
        }
        u32 t[N + 1];
        t[0] = add_cc(ev[0], L);
        ECB_UNROLL
        for (int k = 1; k <= N; k++) t[k] = addc_cc(ev[k], od[k - 1]);
        final_sub(r, t, t[N]);
    }

    ECB_DEV static void sqr(el& r, const el& a) {
        if constexpr (MontKind<P>::kind == 3) {
            u32 T[16];
            sqr_full<N>(T, a.v);
            k256_reduce(r, T);
            return;
        }
        if constexpr (MontKind<P>::kind == 0) {
            u32 T[2 * N];
            sqr_full<N>(T, a.v);
            redc_full(r, T);
            return;
        }
        if constexpr (MontKind<P>::kind == 2) {
            u32 T[25];
            sqr_full<N>(T, a.v);
            T[24] = 0;
            p384_reduce(r, T);
            return;
        }
        if constexpr (MontKind<P>::kind == 1) {
            u32 T[17];
            sqr_full<N>(T, a.v);
            T[16] = 0;
            p256_reduce(r, T);
            return;
        }
        mul(r, a, a);
    }

    // out-of-line copies for the big Weierstrass kernels: one body of each instead of one per call
    // site keeps the window loop inside the instruction cache (ncu: stall_no_instruction was the top
    // stall with everything inlined)
    // (operands and result by value: the ABI then keeps them in registers; by reference they
    // would be forced into local memory)
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

    ECB_DEV static void add(el& r, const el& a, const el& b) {
        u32 t[N];
        u32 c = add_n<N>(t, a.v, b.v);
        if constexpr (LOOSE) {
            refold_carry(r, fold_carry(r, t, c));
            return;
        }
        final_sub(r, t, c);
    }
    ECB_DEV static void sub(el& r, const el& a, const el& b) {
        u32 t[N];
        u32 bw = sub_n<N>(t, a.v, b.v);
        if constexpr (LOOSE) {
            refold_borrow(r, fold_borrow(r, t, bw));
            return;
        }
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
    // constant-time forms for secret operands (ct.cuh).  The loose add / sub above BRANCH on the carry of their fold
    // (once in ~2^32 operand pairs); here the second fold is always executed with that carry as its operand.  The
    // canonical fields are branch-free already (final_sub selects by masks).
    ECB_DEV static void add_ct(el& r, const el& a, const el& b) {
        if constexpr (LOOSE) {
            u32 t[N];
            u32 c = add_n<N>(t, a.v, b.v);
            u32 c2 = fold_carry(r, t, c);
            (void)fold_carry(r, r.v, c2);
            return;
        }
        add(r, a, b);
    }
    ECB_DEV static void sub_ct(el& r, const el& a, const el& b) {
        if constexpr (LOOSE) {
            u32 t[N];
            u32 bw = sub_n<N>(t, a.v, b.v);
            u32 b2 = fold_borrow(r, t, bw);
            (void)fold_borrow(r, r.v, b2);
            return;
        }
        sub(r, a, b);
    }
    ECB_DEV static void neg_ct(el& r, const el& a) {
        el z;
        set_zero(z);
        sub_ct(r, z, a);
    }
    // canonical representative (< p) of a possibly loose value
    ECB_DEV static void canon(el& r, const el& a) {
        if constexpr (LOOSE) {
            final_sub(r, a.v, 0u);
            return;
        }
        copy(r, a);
    }
    ECB_DEV static void dbl(el& r, const el& a) { add(r, a, a); }

    // ---- merged forms for the point formulas.  On the loose fields each is ONE pass with a single fold of the
    // accumulated carries / borrows (a separate add or sub costs 8 + 8 fold instructions and the alu pipe is as
    // loaded as the multiplier in the P-256 kernels); on the canonical fields they are the plain sequences.
    // r = K a, K in {2, 3, 4, 8}
    template <int K>
    ECB_DEV static void mul_small(el& r, const el& a) {
        static_assert(K == 2 || K == 3 || K == 4 || K == 8, "small multiple");
        if constexpr (LOOSE && K != 2) {
            constexpr int S = (K == 8) ? 3 : (K == 4) ? 2 : 1;
            u32 t[N];
            u32 c = a.v[N - 1] >> (32 - S);
            ECB_UNROLL
            for (int i = N - 1; i > 0; i--) t[i] = shl_word(a.v[i - 1], a.v[i], S);
            t[0] = a.v[0] << S;
            if constexpr (K == 3) {
                t[0] = add_cc(t[0], a.v[0]);
                ECB_UNROLL
                for (int i = 1; i < N; i++) t[i] = addc_cc(t[i], a.v[i]);
                c += addc(0u, 0u);
            }
            refold_carry(r, fold_carry_small(r, t, c));
            return;
        }
        if constexpr (K == 2) { add(r, a, a); return; }
        el d;
        add(d, a, a);
        if constexpr (K == 3) { add(r, d, a); return; }
        add(d, d, d);
        if constexpr (K == 4) { copy(r, d); return; }
        add(r, d, d);
    }
    // r = a - b - c
    ECB_DEV static void sub2(el& r, const el& a, const el& b, const el& c) {
        if constexpr (LOOSE) {
            u32 t[N];
            u32 bw = sub_n<N>(t, a.v, b.v);
            bw += sub_n<N>(t, t, c.v);
            refold_borrow(r, fold_borrow_small(r, t, bw));
            return;
        }
        el d;
        sub(d, a, b);
        sub(r, d, c);
    }
    // r = a - b - 2c
    ECB_DEV static void sub_2x(el& r, const el& a, const el& b, const el& c) {
        if constexpr (LOOSE) {
            u32 t[N];
            u32 bw = sub_n<N>(t, a.v, b.v);
            bw += sub_n<N>(t, t, c.v);
            bw += sub_n<N>(t, t, c.v);
            refold_borrow(r, fold_borrow_small(r, t, bw));
            return;
        }
        el d;
        sub(d, a, b);
        sub(d, d, c);
        sub(r, d, c);
    }

    // ---- lazy forms for the doubling (weier.cuh WeiJ::dbl).  On the loose fields the carry out of a fold — the once-in-2^32
    // event that add / sub answer with a second fold — is only COUNTED here (one instruction); the caller looks at the
    // counters once per doubling and, if any is set, recomputes the doubling with the checked forms.  That removes a
    // compare, a branch and a reconvergence point from every field addition.  LZ = false, or a canonical field: the
    // checked forms.
    struct lazy {
        u32 c = 0, b = 0;   // carries counted up, borrows counted down
        ECB_DEV bool any() const { return (c | b) != 0u; }
    };
    template <bool LZ>
    ECB_DEV static void add_z(el& r, const el& a, const el& b, lazy& z) {
        if constexpr (LZ && LOOSE) {
            u32 t[N];
            u32 c = add_n<N>(t, a.v, b.v);
            z.c = fold_carry(r, t, c, &z.c);
            return;
        }
        add(r, a, b);
    }
    template <bool LZ>
    ECB_DEV static void sub_z(el& r, const el& a, const el& b, lazy& z) {
        if constexpr (LZ && LOOSE) {
            u32 t[N];
            u32 bw = sub_n<N>(t, a.v, b.v);
            z.b = fold_borrow(r, t, bw, &z.b);
            return;
        }
        sub(r, a, b);
    }
    template <int K, bool LZ>
    ECB_DEV static void mul_small_z(el& r, const el& a, lazy& z) {
        if constexpr (LZ && LOOSE) {
            if constexpr (K == 2) { add_z<true>(r, a, a, z); return; }
            constexpr int S = (K == 8) ? 3 : (K == 4) ? 2 : 1;
            u32 t[N];
            u32 c = a.v[N - 1] >> (32 - S);
            ECB_UNROLL
            for (int i = N - 1; i > 0; i--) t[i] = shl_word(a.v[i - 1], a.v[i], S);
            t[0] = a.v[0] << S;
            if constexpr (K == 3) {
                t[0] = add_cc(t[0], a.v[0]);
                ECB_UNROLL
                for (int i = 1; i < N; i++) t[i] = addc_cc(t[i], a.v[i]);
                c += addc(0u, 0u);
            }
            z.c = fold_carry_small(r, t, c, &z.c);
            return;
        }
        mul_small<K>(r, a);
    }
    template <bool LZ>
    ECB_DEV static void sub2_z(el& r, const el& a, const el& b, const el& c, lazy& z) {
        if constexpr (LZ && LOOSE) {
            u32 t[N];
            u32 bw = sub_n<N>(t, a.v, b.v);
            bw += sub_n<N>(t, t, c.v);
            z.b = fold_borrow_small(r, t, bw, &z.b);
            return;
        }
        sub2(r, a, b, c);
    }

    ECB_DEV static u32 is_zero(const el& a) {
        u32 o = 0;
        ECB_UNROLL
        for (int i = 0; i < N; i++) o |= a.v[i];
        if constexpr (LOOSE) {  // loose: 0 or p
            u32 q = 0;
            ECB_UNROLL
            for (int i = 0; i < N; i++) q |= a.v[i] ^ P::mod(i);
            return (o == 0 || q == 0) ? 1u : 0u;
        }
        return o == 0 ? 1u : 0u;
    }
    ECB_DEV static u32 eq(const el& a, const el& b) {
        if constexpr (LOOSE) {
            el d;
            sub(d, a, b);
            return is_zero(d);
        }
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
        if constexpr (LOOSE) final_sub(t, t.v, 0u);  // canonical representative
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
    // Montgomery-domain inverse by safegcd (modinv.cuh): (aR)^-1 = a^-1 R^-1, then two products with
    // R^2 bring it back to a^-1 R.  0 -> 0.  This is what the batch-inversion kernels use.
    ECB_DEV static void invert(el& r, const el& a) {
        el c, t, r2;
        canon(c, a);
        u32 p[N];
        ECB_UNROLL
        for (int i = 0; i < N; i++) { p[i] = P::mod(i); r2.v[i] = P::r2(i); }
        sg_modinv<N, (32 * N + 2 + 29) / 30, (N <= 8 ? 22 : 32)>(t.v, c.v, p);
        mul(t, t, r2);
        mul(r, t, r2);
    }
#ifndef ECB_HOSTSIM
    __device__ __forceinline__ static void invert_warp(el& r, const el& a, const u32* jump) {   // one warp, one element (modinv.cuh)
        el c, t, r2;
        canon(c, a);
        u32 p[N];
        ECB_UNROLL
        for (int i = 0; i < N; i++) { p[i] = P::mod(i); r2.v[i] = P::r2(i); }
        sg_modinv_warp<N, (32 * N + 2 + 29) / 30, (N <= 8 ? 22 : 32)>(t.v, c.v, p, jump);
        mul(t, t, r2);
        mul(r, t, r2);
    }
#endif
    ECB_DEV static void invert_fermat(el& r, const el& a) {
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
