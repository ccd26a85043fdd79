/* ecc_oracle.c — CPU restatement of the reference's hot path (vincenthz/eccoxide), plain C.
 *
 * TEST INFRASTRUCTURE AND CPU BASELINE ONLY.  Nothing under eccoxide_b200/ links, loads or calls
 * this file; it is used by tests/ (the checker of the CUDA path), by __graft_entry__.smoke() and by
 * bench.py's cpu_baseline / `--impl reference` legs.  The reference itself is Rust and cannot be
 * built in this image (no cargo/rustc), so this file follows the reference's *algorithms* — same
 * limb radix (64-bit limbs, unsigned __int128 products, 5x51 for 2^255-19, 8x56 for p448,
 * word-by-word Montgomery for the Weierstrass fields), same point formulas, same scalar-mult
 * schedules (4-bit comb with 16-way scan, fixed 4-bit window, w=5 / w=8 wNAF, 256/448-step
 * ladders) — citing file:line for each, so that it can stand in as "the reference's CPU path" when
 * timed.  Parity pinning: tests/test_oracle_golden.py checks every export against the known-answer
 * vectors extracted from the reference's own tests (tests/golden/reference_vectors.json) and
 * against the independent big-integer oracle oracle/pyref.py, libsodium and OpenSSL.
 *
 * Paths are relative to /root/reference.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;
typedef uint64_t u64;
typedef unsigned __int128 u128;

/* =========================================================================================
 * GF(2^255 - 19), 5 x 51-bit limbs  (src/curve/fiat/curve25519_64.rs: carry_mul :217,
 * carry_square :285, carry :351, add :379, sub :400, opp :421, to_bytes :482, from_bytes :631;
 * wrappers src/curve/curve25519.rs:62-117 carry after every add/sub)
 * ========================================================================================= */
typedef struct { u64 v[5]; } fe;
#define M51 ((1ULL << 51) - 1)

static void fe_carry(fe* r, const u64 t[5]) {
    u64 c, a0 = t[0], a1 = t[1], a2 = t[2], a3 = t[3], a4 = t[4];
    c = a0 >> 51; a0 &= M51; a1 += c;
    c = a1 >> 51; a1 &= M51; a2 += c;
    c = a2 >> 51; a2 &= M51; a3 += c;
    c = a3 >> 51; a3 &= M51; a4 += c;
    c = a4 >> 51; a4 &= M51; a0 += 19 * c;
    c = a0 >> 51; a0 &= M51; a1 += c;
    r->v[0] = a0; r->v[1] = a1; r->v[2] = a2; r->v[3] = a3; r->v[4] = a4;
}
static void fe_add(fe* r, const fe* a, const fe* b) {
    u64 t[5];
    for (int i = 0; i < 5; i++) t[i] = a->v[i] + b->v[i];
    fe_carry(r, t);
}
static void fe_sub(fe* r, const fe* a, const fe* b) { /* + 2p bias (curve25519_64.rs:401-405) */
    u64 t[5];
    t[0] = a->v[0] + 0xfffffffffffdaULL - b->v[0];
    for (int i = 1; i < 5; i++) t[i] = a->v[i] + 0xffffffffffffeULL - b->v[i];
    fe_carry(r, t);
}
static void fe_neg(fe* r, const fe* a) {
    fe z = {{0, 0, 0, 0, 0}};
    fe_sub(r, &z, a);
}
static void fe_mul(fe* r, const fe* a, const fe* b) { /* 25 products, 19-fold, carry chain */
    u64 a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3], a4 = a->v[4];
    u64 b0 = b->v[0], b1 = b->v[1], b2 = b->v[2], b3 = b->v[3], b4 = b->v[4];
    u64 b1_19 = 19 * b1, b2_19 = 19 * b2, b3_19 = 19 * b3, b4_19 = 19 * b4;
    u128 t0 = (u128)a0 * b0 + (u128)a1 * b4_19 + (u128)a2 * b3_19 + (u128)a3 * b2_19 + (u128)a4 * b1_19;
    u128 t1 = (u128)a0 * b1 + (u128)a1 * b0 + (u128)a2 * b4_19 + (u128)a3 * b3_19 + (u128)a4 * b2_19;
    u128 t2 = (u128)a0 * b2 + (u128)a1 * b1 + (u128)a2 * b0 + (u128)a3 * b4_19 + (u128)a4 * b3_19;
    u128 t3 = (u128)a0 * b3 + (u128)a1 * b2 + (u128)a2 * b1 + (u128)a3 * b0 + (u128)a4 * b4_19;
    u128 t4 = (u128)a0 * b4 + (u128)a1 * b3 + (u128)a2 * b2 + (u128)a3 * b1 + (u128)a4 * b0;
    u64 c, o[5];
    t1 += (u64)(t0 >> 51); o[0] = (u64)t0 & M51;
    t2 += (u64)(t1 >> 51); o[1] = (u64)t1 & M51;
    t3 += (u64)(t2 >> 51); o[2] = (u64)t2 & M51;
    t4 += (u64)(t3 >> 51); o[3] = (u64)t3 & M51;
    c = (u64)(t4 >> 51);   o[4] = (u64)t4 & M51;
    o[0] += 19 * c;
    c = o[0] >> 51; o[0] &= M51; o[1] += c;
    c = o[1] >> 51; o[1] &= M51; o[2] += c;
    for (int i = 0; i < 5; i++) r->v[i] = o[i];
}
static void fe_sq(fe* r, const fe* a) { /* 15 products */
    u64 a0 = a->v[0], a1 = a->v[1], a2 = a->v[2], a3 = a->v[3], a4 = a->v[4];
    u64 d0 = 2 * a0, d1 = 2 * a1, d2 = 2 * a2, a3_19 = 19 * a3, a4_19 = 19 * a4;
    u128 t0 = (u128)a0 * a0 + (u128)d1 * a4_19 + (u128)d2 * a3_19;
    u128 t1 = (u128)d0 * a1 + (u128)d2 * a4_19 + (u128)a3 * a3_19;
    u128 t2 = (u128)d0 * a2 + (u128)a1 * a1 + (u128)(2 * a3) * a4_19;
    u128 t3 = (u128)d0 * a3 + (u128)d1 * a2 + (u128)a4 * a4_19;
    u128 t4 = (u128)d0 * a4 + (u128)d1 * a3 + (u128)a2 * a2;
    u64 c, o[5];
    t1 += (u64)(t0 >> 51); o[0] = (u64)t0 & M51;
    t2 += (u64)(t1 >> 51); o[1] = (u64)t1 & M51;
    t3 += (u64)(t2 >> 51); o[2] = (u64)t2 & M51;
    t4 += (u64)(t3 >> 51); o[3] = (u64)t3 & M51;
    c = (u64)(t4 >> 51);   o[4] = (u64)t4 & M51;
    o[0] += 19 * c;
    c = o[0] >> 51; o[0] &= M51; o[1] += c;
    c = o[1] >> 51; o[1] &= M51; o[2] += c;
    for (int i = 0; i < 5; i++) r->v[i] = o[i];
}
static void fe_sqn(fe* r, const fe* a, int n) {
    fe_sq(r, a);
    for (int i = 1; i < n; i++) fe_sq(r, r);
}
/* from_bytes_unchecked_le: 255 bits, value may be >= p (curve25519_64.rs:631) */
static void fe_frombytes(fe* r, const u8* s) {
    u64 w[4];
    for (int i = 0; i < 4; i++) {
        w[i] = 0;
        for (int j = 0; j < 8; j++) w[i] |= (u64)s[8 * i + j] << (8 * j);
    }
    r->v[0] = w[0] & M51;
    r->v[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
    r->v[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
    r->v[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
    r->v[4] = (w[3] >> 12) & M51;
}
/* to_bytes: canonical representative (curve25519_64.rs:482) */
static void fe_tobytes(u8* s, const fe* a) {
    u64 t[5], c;
    fe x;
    for (int i = 0; i < 5; i++) t[i] = a->v[i];
    fe_carry(&x, t);
    for (int i = 0; i < 5; i++) t[i] = x.v[i];
    /* q = 1 iff t >= p */
    c = (t[0] + 19) >> 51;
    c = (t[1] + c) >> 51;
    c = (t[2] + c) >> 51;
    c = (t[3] + c) >> 51;
    c = (t[4] + c) >> 51;
    t[0] += 19 * c;
    c = t[0] >> 51; t[0] &= M51; t[1] += c;
    c = t[1] >> 51; t[1] &= M51; t[2] += c;
    c = t[2] >> 51; t[2] &= M51; t[3] += c;
    c = t[3] >> 51; t[3] &= M51; t[4] += c;
    t[4] &= M51;
    u64 w[4];
    w[0] = t[0] | (t[1] << 51);
    w[1] = (t[1] >> 13) | (t[2] << 38);
    w[2] = (t[2] >> 26) | (t[3] << 25);
    w[3] = (t[3] >> 39) | (t[4] << 12);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) s[8 * i + j] = (u8)(w[i] >> (8 * j));
}
static int fe_is_zero(const fe* a) {
    u8 s[32], o = 0;
    fe_tobytes(s, a);
    for (int i = 0; i < 32; i++) o |= s[i];
    return o == 0;
}
static int fe_eq(const fe* a, const fe* b) {
    fe d;
    fe_sub(&d, a, b);
    return fe_is_zero(&d);
}
static void fe_select(fe* r, u64 take_a, const fe* a, const fe* b) {
    u64 m = 0 - (take_a & 1);
    for (int i = 0; i < 5; i++) r->v[i] = (a->v[i] & m) | (b->v[i] & ~m);
}
static void fe_cswap(u64 sw, fe* a, fe* b) {
    u64 m = 0 - (sw & 1);
    for (int i = 0; i < 5; i++) {
        u64 x = (a->v[i] ^ b->v[i]) & m;
        a->v[i] ^= x;
        b->v[i] ^= x;
    }
}
/* pow_2_250_m1 (curve25519.rs:155), invert_or_zero :191, pow_p58 :185 */
static void fe_pow_2_250_m1(fe* t250, fe* z11, const fe* z1) {
    fe z2, z9, t, z5, z10, z20, z50, z100;
    fe_sq(&z2, z1);
    fe_sqn(&t, &z2, 2);
    fe_mul(&z9, &t, z1);
    fe_mul(z11, &z9, &z2);
    fe_sq(&t, z11);
    fe_mul(&z5, &t, &z9);
    fe_sqn(&t, &z5, 5);      fe_mul(&z10, &t, &z5);
    fe_sqn(&t, &z10, 10);    fe_mul(&z20, &t, &z10);
    fe_sqn(&t, &z20, 20);    fe_mul(&t, &t, &z20);
    fe_sqn(&t, &t, 10);      fe_mul(&z50, &t, &z10);
    fe_sqn(&t, &z50, 50);    fe_mul(&z100, &t, &z50);
    fe_sqn(&t, &z100, 100);  fe_mul(&t, &t, &z100);
    fe_sqn(&t, &t, 50);      fe_mul(t250, &t, &z50);
}
static void fe_invert(fe* r, const fe* a) {
    fe t, z11;
    fe_pow_2_250_m1(&t, &z11, a);
    fe_sqn(&t, &t, 5);
    fe_mul(r, &t, &z11);
}
static void fe_pow_p58(fe* r, const fe* a) {
    fe t, z11;
    fe_pow_2_250_m1(&t, &z11, a);
    fe_sqn(&t, &t, 2);
    fe_mul(r, &t, a);
}
static void fe_set_u64(fe* r, u64 x) {
    memset(r, 0, sizeof *r);
    r->v[0] = x & M51;
    r->v[1] = x >> 51;
}

/* =========================================================================================
 * edwards25519 extended points (src/curve/curve25519.rs:592-757)
 * ========================================================================================= */
typedef struct { fe X, Y, Z, T; } ge;
typedef struct { fe ypx, ymx, Z, t2d; } ge_cached; /* CachedPoint, curve25519.rs:994 */

static fe ED_D, ED_D2, ED_SQRTM1, FE_ONE, FE_ZERO;
static ge ED_B, ED_ID;
static ge ED_COMB[64][16];          /* generator_comb(), curve25519.rs:881-902 */
static ge_cached ED_BWNAF[64];      /* GENERATOR_WNAF: odd multiples 1B,3B,..,127B (w = 8), :1080 */
static const u8 L_LE[32] = {0xed, 0xd3, 0xf5, 0x5c, 0x1a, 0x63, 0x12, 0x58, 0xd6, 0x9c, 0xf7, 0xa2, 0xde, 0xf9, 0xde, 0x14,
                            0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x10};

static void ge_double(ge* r, const ge* p) { /* double_parts :604 + double :669 — 4 M + 4 S */
    fe a, b, c, d, xy, e, g, f, h, t;
    fe_sq(&a, &p->X);
    fe_sq(&b, &p->Y);
    fe_sq(&c, &p->Z);
    fe_add(&c, &c, &c);
    fe_neg(&d, &a);
    fe_add(&xy, &p->X, &p->Y);
    fe_sq(&e, &xy);
    fe_add(&t, &a, &b);
    fe_sub(&e, &e, &t);
    fe_add(&g, &d, &b);
    fe_sub(&f, &g, &c);
    fe_sub(&h, &d, &b);
    fe_mul(&r->X, &e, &f);
    fe_mul(&r->Y, &g, &h);
    fe_mul(&r->Z, &f, &g);
    fe_mul(&r->T, &e, &h);
}
static void ge_add(ge* r, const ge* p, const ge* q) { /* Point::add :695 — 9 M */
    fe aa, bb, cc, dd, e, f, g, h, t1, t2;
    fe_sub(&t1, &p->Y, &p->X);
    fe_sub(&t2, &q->Y, &q->X);
    fe_mul(&aa, &t1, &t2);
    fe_add(&t1, &p->Y, &p->X);
    fe_add(&t2, &q->Y, &q->X);
    fe_mul(&bb, &t1, &t2);
    fe_mul(&cc, &ED_D2, &p->T);
    fe_mul(&cc, &cc, &q->T);
    fe_mul(&dd, &p->Z, &q->Z);
    fe_add(&dd, &dd, &dd);
    fe_sub(&e, &bb, &aa);
    fe_sub(&f, &dd, &cc);
    fe_add(&g, &dd, &cc);
    fe_add(&h, &bb, &aa);
    fe_mul(&r->X, &e, &f);
    fe_mul(&r->Y, &g, &h);
    fe_mul(&r->Z, &f, &g);
    fe_mul(&r->T, &e, &h);
}
static void ge_to_cached(ge_cached* c, const ge* p) {
    fe_add(&c->ypx, &p->Y, &p->X);
    fe_sub(&c->ymx, &p->Y, &p->X);
    c->Z = p->Z;
    fe_mul(&c->t2d, &p->T, &ED_D2);
}
static void ge_add_cached(ge* r, const ge* p, const ge_cached* q, int negate) { /* add_cached :715 — 8 M */
    fe aa, bb, cc, dd, e, f, g, h, t;
    fe_sub(&t, &p->Y, &p->X);
    fe_mul(&aa, &t, negate ? &q->ypx : &q->ymx);
    fe_add(&t, &p->Y, &p->X);
    fe_mul(&bb, &t, negate ? &q->ymx : &q->ypx);
    fe_mul(&cc, &p->T, &q->t2d);
    if (negate) fe_neg(&cc, &cc);
    fe_mul(&dd, &p->Z, &q->Z);
    fe_add(&dd, &dd, &dd);
    fe_sub(&e, &bb, &aa);
    fe_sub(&f, &dd, &cc);
    fe_add(&g, &dd, &cc);
    fe_add(&h, &bb, &aa);
    fe_mul(&r->X, &e, &f);
    fe_mul(&r->Y, &g, &h);
    fe_mul(&r->Z, &f, &g);
    fe_mul(&r->T, &e, &h);
}
static void ge_select(ge* r, u64 take_a, const ge* a, const ge* b) { /* ct_select :1186 */
    fe_select(&r->X, take_a, &a->X, &b->X);
    fe_select(&r->Y, take_a, &a->Y, &b->Y);
    fe_select(&r->Z, take_a, &a->Z, &b->Z);
    fe_select(&r->T, take_a, &a->T, &b->T);
}
static void ge_from_affine(ge* r, const fe* x, const fe* y) {
    r->X = *x;
    r->Y = *y;
    r->Z = FE_ONE;
    fe_mul(&r->T, x, y);
}
static void ge_to_affine_bytes(u8* out64, const ge* p) { /* to_affine :663 + to_bytes_le */
    fe zi, x, y;
    fe_invert(&zi, &p->Z);
    fe_mul(&x, &p->X, &zi);
    fe_mul(&y, &p->Y, &zi);
    fe_tobytes(out64, &x);
    fe_tobytes(out64 + 32, &y);
}
/* Point::from_coordinate :649 */
static int ge_on_curve(const fe* x, const fe* y) {
    fe xx, yy, l, r;
    fe_sq(&xx, x);
    fe_sq(&yy, y);
    fe_sub(&l, &yy, &xx);
    fe_mul(&r, &xx, &yy);
    fe_mul(&r, &r, &ED_D);
    fe_add(&r, &r, &FE_ONE);
    return fe_eq(&l, &r);
}
static int bytes_lt_le(const u8* a, const u8* m, int n) { /* a < m, little-endian */
    for (int i = n - 1; i >= 0; i--) {
        if (a[i] < m[i]) return 1;
        if (a[i] > m[i]) return 0;
    }
    return 0;
}
static int fe_bytes_canonical(const u8* s) { /* value (all 256 bits) < p */
    static const u8 P_LE[32] = {0xed, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff,
                                0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0x7f};
    return bytes_lt_le(s, P_LE, 32);
}
/* Point::mul_base :840-869: 64 windows, nibble i (little-endian) selects from window i by a full scan */
static void ed_mul_base(ge* r, const u8* k_le) {
    ge q = ED_ID, sel;
    for (int i = 0; i < 64; i++) {
        unsigned digit = (i & 1) ? (k_le[i >> 1] >> 4) : (k_le[i >> 1] & 15u);
        sel = ED_ID;
        for (unsigned j = 0; j < 16; j++) ge_select(&sel, j == digit, &ED_COMB[i][j], &sel);
        ge_add(&q, &q, &sel);
    }
    *r = q;
}
/* Point::scale_bytes :746: 256 x (double, add, select), MSB first */
static void ed_scale(ge* r, const ge* p, const u8* k_le) {
    ge q = ED_ID, added;
    for (int bit = 255; bit >= 0; bit--) {
        ge_double(&q, &q);
        ge_add(&added, &q, p);
        ge_select(&q, (k_le[bit >> 3] >> (bit & 7)) & 1, &added, &q);
    }
    *r = q;
}
/* wnaf :938 — digits of a little-endian 256-bit scalar, width w */
static int ed_wnaf(signed char* naf, const u8* k_le, int w) {
    u64 k[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 32; i++) k[i / 8] |= (u64)k_le[i] << (8 * (i % 8));
    int cnt = 0;
    const long width = 1L << w, half = 1L << (w - 1);
    while (k[0] | k[1] | k[2] | k[3] | k[4]) {
        long digit = 0;
        if (k[0] & 1) {
            long m = (long)(k[0] & (u64)(width - 1));
            digit = m >= half ? m - width : m;
            if (digit > 0) {
                u64 bw = (u64)digit;
                for (int i = 0; i < 5 && bw; i++) { u64 o = k[i]; k[i] = o - bw; bw = o < bw; }
            } else {
                u64 c = (u64)(-digit);
                for (int i = 0; i < 5 && c; i++) { u64 o = k[i]; k[i] = o + c; c = k[i] < o; }
            }
        }
        naf[cnt++] = (signed char)digit;
        for (int i = 0; i < 4; i++) k[i] = (k[i] >> 1) | (k[i + 1] << 63);
        k[4] >>= 1;
    }
    return cnt;
}
/* double_scalar_mul_base_vartime :1157: s*B + k*P, w = 8 for B (static table), w = 5 for P */
static void ed_double_scalar_vartime(ge* r, const u8* s_le, const u8* k_le, const ge* p) {
    signed char nb[260], np[260];
    int cb = ed_wnaf(nb, s_le, 8), cp = ed_wnaf(np, k_le, 5);
    ge_cached tp[8];
    ge p2, t = *p;
    ge_double(&p2, p);
    ge_cached c2;
    ge_to_cached(&c2, &p2);
    for (int i = 0; i < 8; i++) { /* odd_multiples :1061 */
        ge_to_cached(&tp[i], &t);
        if (i < 7) ge_add_cached(&t, &t, &c2, 0);
    }
    ge q = ED_ID;
    int top = cb > cp ? cb : cp;
    for (int i = top - 1; i >= 0; i--) {
        ge_double(&q, &q);
        int d = i < cb ? nb[i] : 0;
        if (d > 0) ge_add_cached(&q, &q, &ED_BWNAF[d >> 1], 0);
        else if (d < 0) ge_add_cached(&q, &q, &ED_BWNAF[(-d) >> 1], 1);
        d = i < cp ? np[i] : 0;
        if (d > 0) ge_add_cached(&q, &q, &tp[d >> 1], 0);
        else if (d < 0) ge_add_cached(&q, &q, &tp[(-d) >> 1], 1);
    }
    *r = q;
}
/* decode_point (protocol/ed25519.rs:38-59) -> decompress (curve25519.rs:772) -> sqrt_div :258 */
static int ed_decode(ge* r, const u8* enc) {
    u8 yb[32];
    memcpy(yb, enc, 32);
    unsigned sign = yb[31] >> 7;
    yb[31] &= 0x7f;
    if (!fe_bytes_canonical(yb)) return 0;
    fe y, yy, u, v, v3, v7, t, x, chk, nu;
    fe_frombytes(&y, yb);
    fe_sq(&yy, &y);
    fe_sub(&u, &yy, &FE_ONE);
    fe_mul(&v, &yy, &ED_D);
    fe_add(&v, &v, &FE_ONE);
    fe_sq(&v3, &v); fe_mul(&v3, &v3, &v);
    fe_sq(&v7, &v3); fe_mul(&v7, &v7, &v);
    fe_mul(&t, &u, &v7);
    fe_pow_p58(&t, &t);
    fe_mul(&x, &u, &v3);
    fe_mul(&x, &x, &t);
    fe_sq(&chk, &x);
    fe_mul(&chk, &chk, &v);
    fe_neg(&nu, &u);
    if (fe_eq(&chk, &u)) {
    } else if (fe_eq(&chk, &nu)) {
        fe_mul(&x, &x, &ED_SQRTM1);
    } else {
        return 0;
    }
    u8 xb[32];
    fe_tobytes(xb, &x);
    int xzero = 1;
    for (int i = 0; i < 32; i++) xzero &= xb[i] == 0;
    if (xzero && sign) return 0;
    if ((unsigned)(xb[0] & 1) != sign) fe_neg(&x, &x);
    ge_from_affine(r, &x, &y);
    return 1;
}
/* verify (protocol/ed25519.rs:119-147), k = reduce_wide_le(SHA-512(R||A||M)) supplied */
static int ed_verify_prehashed(const u8* a_enc, const u8* r_enc, const u8* s_le, const u8* k_le) {
    ge A, R, lhs;
    if (!ed_decode(&A, a_enc)) return 0;
    if (!ed_decode(&R, r_enc)) return 0;
    if (!bytes_lt_le(s_le, L_LE, 32)) return 0;
    if (!bytes_lt_le(k_le, L_LE, 32)) return 0;
    fe_neg(&A.X, &A.X);
    fe_neg(&A.T, &A.T);
    ed_double_scalar_vartime(&lhs, s_le, k_le, &A);
    /* Point::eq :1200 — projective, R has Z = 1 */
    fe t1, t2;
    fe_mul(&t1, &R.X, &lhs.Z);
    fe_mul(&t2, &R.Y, &lhs.Z);
    return fe_eq(&t1, &lhs.X) && fe_eq(&t2, &lhs.Y);
}

/* X25519 (protocol/x25519.rs:36-45; ladder curve25519.rs:474-513) */
static void x25519_one(u8* out, const u8* scalar, const u8* ubytes) {
    u8 k[32], ub[32];
    memcpy(k, scalar, 32);
    k[0] &= 248; k[31] &= 127; k[31] |= 64;       /* clamp :15 */
    memcpy(ub, ubytes, 32);
    ub[31] &= 0x7f;                                /* decode_u :24 */
    fe x1, x2, z2, x3, z3, a24;
    fe_frombytes(&x1, ub);
    {   /* MontgomeryPoint::scale_bytes calls self.u() first (:531-536): one inversion of z = 1 */
        fe one = FE_ONE, zi;
        fe_invert(&zi, &one);
        fe_mul(&x1, &x1, &zi);
    }
    fe_set_u64(&a24, 121666);
    x2 = FE_ONE; z2 = FE_ZERO; x3 = x1; z3 = FE_ONE;
    u64 swap = 0;
    for (int t = 255; t >= 0; t--) {
        u64 bit = (k[t >> 3] >> (t & 7)) & 1;
        swap ^= bit;
        fe_cswap(swap, &x2, &x3);
        fe_cswap(swap, &z2, &z3);
        swap = bit;
        fe a, aa, b, bb, e, c, d, da, cb, s;
        fe_add(&a, &x2, &z2);  fe_sq(&aa, &a);
        fe_sub(&b, &x2, &z2);  fe_sq(&bb, &b);
        fe_sub(&e, &aa, &bb);
        fe_add(&c, &x3, &z3);  fe_sub(&d, &x3, &z3);
        fe_mul(&da, &d, &a);   fe_mul(&cb, &c, &b);
        fe_add(&s, &da, &cb);  fe_sq(&x3, &s);
        fe_sub(&s, &da, &cb);  fe_sq(&s, &s);  fe_mul(&z3, &x1, &s);
        fe_mul(&x2, &aa, &bb);
        fe_mul(&s, &a24, &e);  /* full field multiplication, :505 */
        fe_add(&s, &bb, &s);
        fe_mul(&z2, &e, &s);
    }
    fe_cswap(swap, &x2, &x3);
    fe_cswap(swap, &z2, &z3);
    fe zi, r;
    fe_invert(&zi, &z2); /* invert_or_zero: 0 -> 0 */
    fe_mul(&r, &x2, &zi);
    fe_tobytes(out, &r);
}

/* =========================================================================================
 * GF(2^448 - 2^224 - 1), 8 x 56-bit limbs (src/curve/fiat/p448_solinas_64.rs; curve448.rs:45-174)
 * and X448 (curve448.rs:263-302, protocol/x448.rs:16-45)
 * ========================================================================================= */
typedef struct { u64 v[8]; } fe448;
#define M56 ((1ULL << 56) - 1)
static void f448_carry(fe448* r, u64 t[8]) {
    for (int rep = 0; rep < 2; rep++) {
        u64 c = 0;
        for (int i = 0; i < 8; i++) { t[i] += c; c = t[i] >> 56; t[i] &= M56; }
        t[0] += c;
        t[4] += c;
    }
    for (int i = 0; i < 8; i++) r->v[i] = t[i];
}
static void f448_add(fe448* r, const fe448* a, const fe448* b) {
    u64 t[8];
    for (int i = 0; i < 8; i++) t[i] = a->v[i] + b->v[i];
    f448_carry(r, t);
}
static void f448_sub(fe448* r, const fe448* a, const fe448* b) { /* + 2p */
    u64 t[8];
    for (int i = 0; i < 8; i++) t[i] = a->v[i] + ((i == 4) ? 2 * (M56 - 1) : 2 * M56) - b->v[i];
    f448_carry(r, t);
}
static void f448_mul(fe448* r, const fe448* a, const fe448* b) {
    u128 t[16];
    for (int i = 0; i < 16; i++) t[i] = 0;
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) t[i + j] += (u128)a->v[i] * b->v[j];
    for (int k = 14; k >= 8; k--) { t[k - 8] += t[k]; t[k - 4] += t[k]; } /* 2^448 = 2^224 + 1 */
    u64 o[8];
    u128 c = 0;
    for (int i = 0; i < 8; i++) { c += t[i]; o[i] = (u64)c & M56; c >>= 56; }
    /* c * 2^448 = c * (2^224 + 1); c < 2^70: split it over two limbs */
    u64 clo = (u64)c & M56, chi = (u64)(c >> 56);
    o[0] += clo; o[1] += chi; o[4] += clo; o[5] += chi;
    f448_carry(r, o);
}
static void f448_sq(fe448* r, const fe448* a) { f448_mul(r, a, a); }
static void f448_sqn(fe448* r, const fe448* a, int n) {
    f448_sq(r, a);
    for (int i = 1; i < n; i++) f448_sq(r, r);
}
static void f448_frombytes(fe448* r, const u8* s) {
    for (int i = 0; i < 8; i++) {
        r->v[i] = 0;
        for (int j = 0; j < 7; j++) r->v[i] |= (u64)s[7 * i + j] << (8 * j);
    }
}
static void f448_tobytes(u8* s, const fe448* a) {
    u64 t[8];
    fe448 x;
    for (int i = 0; i < 8; i++) t[i] = a->v[i];
    f448_carry(&x, t);
    for (;;) { /* strict limbs: every limb < 2^56 */
        int loose = 0;
        for (int i = 0; i < 8; i++) loose |= (x.v[i] >> 56) != 0;
        if (!loose) break;
        f448_carry(&x, x.v);
    }
    /* subtract p if x >= p: x + 2^224 + 1 overflows 2^448 */
    u64 s2[8], c = 1;
    for (int i = 0; i < 8; i++) { s2[i] = x.v[i] + c + (i == 4 ? 1 : 0); c = s2[i] >> 56; s2[i] &= M56; }
    const u64* o = c ? s2 : x.v;
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 7; j++) s[7 * i + j] = (u8)(o[i] >> (8 * j));
}
static void f448_cswap(u64 sw, fe448* a, fe448* b) {
    u64 m = 0 - (sw & 1);
    for (int i = 0; i < 8; i++) {
        u64 x = (a->v[i] ^ b->v[i]) & m;
        a->v[i] ^= x;
        b->v[i] ^= x;
    }
}
static void f448_invert(fe448* r, const fe448* a) { /* a^(p-2), p-2 = 2^448 - 2^224 - 3 */
    fe448 x2, x3, x6, x12, x24, x27, x54, x108, x111, x222, x223, t;
    f448_sq(&t, a);           f448_mul(&x2, &t, a);
    f448_sq(&t, &x2);         f448_mul(&x3, &t, a);
    f448_sqn(&t, &x3, 3);     f448_mul(&x6, &t, &x3);
    f448_sqn(&t, &x6, 6);     f448_mul(&x12, &t, &x6);
    f448_sqn(&t, &x12, 12);   f448_mul(&x24, &t, &x12);
    f448_sqn(&t, &x24, 3);    f448_mul(&x27, &t, &x3);
    f448_sqn(&t, &x27, 27);   f448_mul(&x54, &t, &x27);
    f448_sqn(&t, &x54, 54);   f448_mul(&x108, &t, &x54);
    f448_sqn(&t, &x108, 3);   f448_mul(&x111, &t, &x3);
    f448_sqn(&t, &x111, 111); f448_mul(&x222, &t, &x111);
    f448_sq(&t, &x222);       f448_mul(&x223, &t, a);
    f448_sqn(&t, &x223, 223); f448_mul(&t, &t, &x222);
    f448_sqn(&t, &t, 2);      f448_mul(r, &t, a);
}
static void x448_one(u8* out, const u8* scalar, const u8* ubytes) {
    u8 k[56];
    memcpy(k, scalar, 56);
    k[0] &= 252; k[55] |= 128;
    fe448 x1, x2, z2, x3, z3, a24, one, zero;
    memset(&one, 0, sizeof one); one.v[0] = 1;
    memset(&zero, 0, sizeof zero);
    memset(&a24, 0, sizeof a24); a24.v[0] = 39082;
    f448_frombytes(&x1, ubytes);
    x2 = one; z2 = zero; x3 = x1; z3 = one;
    u64 swap = 0;
    for (int t = 447; t >= 0; t--) {
        u64 bit = (k[t >> 3] >> (t & 7)) & 1;
        swap ^= bit;
        f448_cswap(swap, &x2, &x3);
        f448_cswap(swap, &z2, &z3);
        swap = bit;
        fe448 a, aa, b, bb, e, c, d, da, cb, s;
        f448_add(&a, &x2, &z2);  f448_sq(&aa, &a);
        f448_sub(&b, &x2, &z2);  f448_sq(&bb, &b);
        f448_sub(&e, &aa, &bb);
        f448_add(&c, &x3, &z3);  f448_sub(&d, &x3, &z3);
        f448_mul(&da, &d, &a);   f448_mul(&cb, &c, &b);
        f448_add(&s, &da, &cb);  f448_sq(&x3, &s);
        f448_sub(&s, &da, &cb);  f448_sq(&s, &s);  f448_mul(&z3, &x1, &s);
        f448_mul(&x2, &aa, &bb);
        f448_mul(&s, &a24, &e);
        f448_add(&s, &bb, &s);
        f448_mul(&z2, &e, &s);
    }
    f448_cswap(swap, &x2, &x3);
    f448_cswap(swap, &z2, &z3);
    fe448 zi, r;
    f448_invert(&zi, &z2);
    f448_mul(&r, &x2, &zi);
    f448_tobytes(out, &r);
}

/* =========================================================================================
 * Weierstrass curves: Montgomery fields on 4 / 6 limbs
 * ========================================================================================= */
#define NL 4
#include "mont_tmpl.h"
#undef NL
#define NL 6
#include "mont_tmpl.h"
#undef NL

static curve4 P256, K256;   /* K256 = secp256k1, the reference's p256k1 (src/curve/sec2/p256k1.rs; a = 0 path) */
static curve6 P384, BLSG1;

static void hex2be(u8* out, const char* hex, int len) {
    for (int i = 0; i < len; i++) {
        unsigned v = 0;
        for (int j = 0; j < 2; j++) {
            char ch = hex[2 * i + j];
            v = v * 16 + (unsigned)(ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10);
        }
        out[i] = (u8)v;
    }
}

/* ECDSA verify_hashed (protocol/ecdsa.rs:205-222) for a curve whose scalar field has the same limb
 * count as its base field (p256r1, p384r1).  Generated per NL by the macro below. */
#define DEFINE_ECDSA(NLV)                                                                                          \
    static int ecdsa_verify##NLV(const curve##NLV* c, const u8* q_xy, const u8* z_be, const u8* rs_be, int* bad) { \
        const field##NLV* FN = &c->fn;                                                                             \
        int sb = c->sbytes;                                                                                        \
        pt##NLV Q;                                                                                                 \
        *bad = 0;                                                                                                  \
        if (!pt_from_xy_be##NLV(c, &Q, q_xy)) { *bad = 1; return 0; }                                              \
        fe##NLV r, s, z, si, u1, u2;                                                                               \
        u64 rr[NLV], ss[NLV], zz[NLV];                                                                             \
        raw_from_be##NLV(rr, rs_be, sb);                                                                           \
        raw_from_be##NLV(ss, rs_be + sb, sb);                                                                      \
        raw_from_be##NLV(zz, z_be, sb);                                                                            \
        int rz = 1, sz = 1;                                                                                        \
        for (int i = 0; i < NLV; i++) { rz &= rr[i] == 0; sz &= ss[i] == 0; }                                      \
        if (rz || sz || ge_raw##NLV(rr, FN->p) || ge_raw##NLV(ss, FN->p)) return 0; /* Signature::from_bytes */    \
        while (ge_raw##NLV(zz, FN->p)) sub_raw##NLV(zz, zz, FN->p);              /* digest_to_scalar :340 */       \
        fe##NLV t;                                                                                                 \
        memcpy(t.v, rr, sizeof rr); f_mul##NLV(FN, &r, &t, &FN->r2);                                               \
        memcpy(t.v, ss, sizeof ss); f_mul##NLV(FN, &s, &t, &FN->r2);                                               \
        memcpy(t.v, zz, sizeof zz); f_mul##NLV(FN, &z, &t, &FN->r2);                                               \
        f_inv##NLV(FN, &si, &s);                                   /* Scalar::inverse, p256r1.rs:117 */            \
        f_mul##NLV(FN, &u1, &z, &si);                                                                              \
        f_mul##NLV(FN, &u2, &r, &si);                                                                              \
        u8 k1[8 * NLV], k2[8 * NLV];                                                                               \
        f_to_be##NLV(FN, k1, &u1, sb);                                                                             \
        f_to_be##NLV(FN, k2, &u2, sb);                                                                             \
        pt##NLV A, B, R;                                                                                           \
        pt_mul_base##NLV(c, &A, k1, sb);                           /* Point::mul_base (comb) */                    \
        pt_mul_wnaf##NLV(c, &B, &Q, k2, sb);                       /* mul_vartime (w = 5 wNAF) */                  \
        pt_add##NLV(c, &R, &A, &B);                                                                                \
        fe##NLV x;                                                                                                 \
        if (!pt_to_affine##NLV(c, &x, 0, &R)) return 0;            /* identity => reject :218-221 */               \
        u64 xw[NLV];                                                                                               \
        f_to_raw##NLV(&c->fp, xw, &x);                                                                             \
        while (ge_raw##NLV(xw, FN->p)) sub_raw##NLV(xw, xw, FN->p); /* field_to_scalar :363 */                     \
        for (int i = 0; i < NLV; i++)                                                                              \
            if (xw[i] != rr[i]) return 0;                                                                          \
        return 1;                                                                                                  \
    }
DEFINE_ECDSA(4)
DEFINE_ECDSA(6)

/* =========================================================================================
 * init
 * ========================================================================================= */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void bls_extra_init(void);
static void do_init(void) {
    u8 b[32];
    fe_set_u64(&FE_ONE, 1);
    fe_set_u64(&FE_ZERO, 0);
    /* d = -121665/121666, sqrt(-1) = 2^((p-1)/4): computed, not embedded (curve25519.rs:395-423) */
    fe n, dd, t;
    fe_set_u64(&n, 121665);
    fe_set_u64(&dd, 121666);
    fe_invert(&t, &dd);
    fe_mul(&ED_D, &n, &t);
    fe_neg(&ED_D, &ED_D);
    fe_add(&ED_D2, &ED_D, &ED_D);
    {   /* 2^((p-1)/4) = 2^(2^253 - 5): (2^((p-5)/8))^2 * 2 ... use pow_p58: 2^((p-5)/8) = 2^(2^252-3) */
        fe two, e;
        fe_set_u64(&two, 2);
        fe_pow_p58(&e, &two);       /* 2^(2^252 - 3) */
        fe_sq(&e, &e);              /* 2^(2^253 - 6) */
        fe_mul(&ED_SQRTM1, &e, &two); /* 2^(2^253 - 5) */
    }
    /* base point: y = 4/5, x = the even root (RFC 8032 §5.1) */
    {
        fe four, five, y;
        fe_set_u64(&four, 4);
        fe_set_u64(&five, 5);
        fe_invert(&t, &five);
        fe_mul(&y, &four, &t);
        fe_tobytes(b, &y);
        ED_ID.X = FE_ZERO; ED_ID.Y = FE_ONE; ED_ID.Z = FE_ONE; ED_ID.T = FE_ZERO;
        ed_decode(&ED_B, b);
    }
    /* comb: [i][j] = j * 16^i * B, affine (Z = 1); [i][0] = identity (params/comb/curve25519.rs:7-8) */
    {
        ge base = ED_B;
        for (int i = 0; i < 64; i++) {
            ED_COMB[i][0] = ED_ID;
            ED_COMB[i][1] = base;
            for (int j = 2; j < 16; j++) ge_add(&ED_COMB[i][j], &ED_COMB[i][j - 1], &base);
            ge nb;
            ge_add(&nb, &ED_COMB[i][15], &base);
            for (int j = 1; j < 16; j++) {
                u8 xy[64];
                fe x, y;
                ge_to_affine_bytes(xy, &ED_COMB[i][j]);
                fe_frombytes(&x, xy);
                fe_frombytes(&y, xy + 32);
                ge_from_affine(&ED_COMB[i][j], &x, &y);
            }
            u8 xy[64];
            fe x, y;
            ge_to_affine_bytes(xy, &nb);
            fe_frombytes(&x, xy);
            fe_frombytes(&y, xy + 32);
            ge_from_affine(&base, &x, &y);
        }
    }
    /* GENERATOR_WNAF: (2i+1) * B cached, i = 0..63 (params/comb/curve25519.rs:1100-1107) */
    {
        ge b2, tt = ED_B;
        ge_double(&b2, &ED_B);
        for (int i = 0; i < 64; i++) {
            u8 xy[64];
            fe x, y;
            ge a;
            ge_to_affine_bytes(xy, &tt);
            fe_frombytes(&x, xy);
            fe_frombytes(&y, xy + 32);
            ge_from_affine(&a, &x, &y);
            ge_to_cached(&ED_BWNAF[i], &a);
            ge_add(&tt, &tt, &b2);
        }
    }
    /* Weierstrass domain parameters (SEC 2 / FIPS 186-4 / BLS12-381; the reference keeps the same
     * values in src/params/sec2.rs:1041- and src/params/bls12_381.rs:25-85) */
    u8 p[48], nn[48], bb[48], gx[48], gy[48];
    hex2be(p, "ffffffff00000001000000000000000000000000ffffffffffffffffffffffff", 32);
    hex2be(nn, "ffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551", 32);
    hex2be(bb, "5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b", 32);
    hex2be(gx, "6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296", 32);
    hex2be(gy, "4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5", 32);
    curve_init4(&P256, p, nn, bb, gx, gy, 32, 32, 0);
    hex2be(p, "fffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffffeffffffff0000000000000000ffffffff", 48);
    hex2be(nn, "ffffffffffffffffffffffffffffffffffffffffffffffffc7634d81f4372ddf581a0db248b0a77aecec196accc52973", 48);
    hex2be(bb, "b3312fa7e23ee7e4988e056be3f82d19181d9c6efe8141120314088f5013875ac656398d8a2ed19d2a85c8edd3ec2aef", 48);
    hex2be(gx, "aa87ca22be8b05378eb1c71ef320ad746e1d3b628ba79b9859f741e082542a385502f25dbf55296c3a545e3872760ab7", 48);
    hex2be(gy, "3617de4a96262c6f5d9e98bf9292dc29f8f41dbd289a147ce9da3113b5f0b8c00a60b1ce1d7e819d7a431d7c90ea0e5f", 48);
    curve_init6(&P384, p, nn, bb, gx, gy, 48, 48, 0);
    hex2be(p, "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab", 48);
    hex2be(nn, "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001", 32);
    memset(bb, 0, 48); bb[47] = 4;
    hex2be(gx, "17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb", 48);
    hex2be(gy, "08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1", 48);
    curve_init6(&BLSG1, p, nn, bb, gx, gy, 48, 32, 1);
    /* secp256k1 (SEC 2 v2 2.4.1; src/params/sec2.rs:908-) */
    hex2be(p, "fffffffffffffffffffffffffffffffffffffffffffffffffffffffefffffc2f", 32);
    hex2be(nn, "fffffffffffffffffffffffffffffffebaaedce6af48a03bbfd25e8cd0364141", 32);
    memset(bb, 0, 48); bb[31] = 7;
    hex2be(gx, "79be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798", 32);
    hex2be(gy, "483ada7726a3c4655da4fbfc0e1108a8fd17b448a68554199c47d08ffb10d4b8", 32);
    curve_init4(&K256, p, nn, bb, gx, gy, 32, 32, 1);
    bls_extra_init();
}
void orc_init(void) { pthread_once(&g_once, do_init); }

/* =========================================================================================
 * batch drivers: static contiguous partition over `nthreads` pthreads (the reference has no
 * threading; this is the "all host cores" arm of the CPU baseline)
 * ========================================================================================= */
typedef struct job {
    void (*fn)(struct job*, size_t lo, size_t hi);
    size_t lo, hi;
    int curve, mode;
    const u8 *a, *b, *c, *d;
    u8 *o, *o2;
    long bad;   /* first bad index in this slice, -1 if none */
    int code;
} job;
static void* job_main(void* p) {
    job* j = (job*)p;
    j->fn(j, j->lo, j->hi);
    return 0;
}
static long run_jobs(job* tmpl, size_t n, int nthreads, int* code) {
    orc_init();
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
    job* js = (job*)malloc(sizeof(job) * (size_t)nthreads);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int t = 0; t < nthreads; t++) {
        js[t] = *tmpl;
        js[t].lo = n * (size_t)t / (size_t)nthreads;
        js[t].hi = n * (size_t)(t + 1) / (size_t)nthreads;
        js[t].bad = -1;
        js[t].code = 0;
    }
    for (int t = 1; t < nthreads; t++) pthread_create(&th[t], 0, job_main, &js[t]);
    job_main(&js[0]);
    for (int t = 1; t < nthreads; t++) pthread_join(th[t], 0);
    long bad = -1;
    for (int t = 0; t < nthreads; t++)
        if (js[t].bad >= 0 && (bad < 0 || js[t].bad < bad)) { bad = js[t].bad; if (code) *code = js[t].code; }
    free(js);
    free(th);
    return bad;
}
#define BAD(j, i, c) do { if ((j)->bad < 0) { (j)->bad = (long)(i); (j)->code = (c); } } while (0)
enum { ORC_NONCANONICAL_SCALAR = 1, ORC_BAD_POINT = 2 };

static void j_ed_mul_base(job* j, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
        const u8* k = j->a + 32 * i;
        if (!bytes_lt_le(k, L_LE, 32)) { BAD(j, i, ORC_NONCANONICAL_SCALAR); memset(j->o + 64 * i, 0, 64); continue; }
        ge r;
        ed_mul_base(&r, k);
        ge_to_affine_bytes(j->o + 64 * i, &r);
    }
}
static void j_ed_mul(job* j, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
        const u8 *k = j->a + 32 * i, *xy = j->b + 64 * i;
        if (!bytes_lt_le(k, L_LE, 32)) { BAD(j, i, ORC_NONCANONICAL_SCALAR); memset(j->o + 64 * i, 0, 64); continue; }
        fe x, y;
        fe_frombytes(&x, xy);
        fe_frombytes(&y, xy + 32);
        if (!fe_bytes_canonical(xy) || !fe_bytes_canonical(xy + 32) || !ge_on_curve(&x, &y)) {
            BAD(j, i, ORC_BAD_POINT); memset(j->o + 64 * i, 0, 64); continue;
        }
        ge p, r;
        ge_from_affine(&p, &x, &y);
        ed_scale(&r, &p, k);
        ge_to_affine_bytes(j->o + 64 * i, &r);
    }
}
static void j_ed_verify(job* j, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) j->o[i] = (u8)ed_verify_prehashed(j->a + 32 * i, j->b + 32 * i, j->c + 32 * i, j->d + 32 * i);
}
static void j_x25519(job* j, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) x25519_one(j->o + 32 * i, j->a + 32 * i, j->b + 32 * i);
}
static void j_x448(job* j, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) x448_one(j->o + 56 * i, j->a + 56 * i, j->b + 56 * i);
}

#define WEI_BODY(NLV, CURVE, MODE)                                                                         \
    {                                                                                                      \
        const curve##NLV* c = (CURVE);                                                                     \
        int fb = c->fbytes, sb = c->sbytes;                                                                \
        for (size_t i = lo; i < hi; i++) {                                                                 \
            const u8* k = j->a + (size_t)sb * i;                                                           \
            u8* out = j->o + (size_t)2 * fb * i;                                                           \
            u64 kw[NLV];                                                                                   \
            raw_from_be##NLV(kw, k, sb);                                                                   \
            memset(out, 0, (size_t)2 * fb);                                                                \
            if (j->o2) j->o2[i] = 0;                                                                       \
            if (ge_raw##NLV(kw, c->fn.p)) { BAD(j, i, ORC_NONCANONICAL_SCALAR); continue; }                \
            pt##NLV P, R;                                                                                  \
            if ((MODE) != 1) {                                                                             \
                if (j->c && j->c[i]) pt_inf##NLV(c, &P);                                                   \
                else if (!pt_from_xy_be##NLV(c, &P, j->b + (size_t)2 * fb * i)) { BAD(j, i, ORC_BAD_POINT); continue; } \
            }                                                                                              \
            if ((MODE) == 0) pt_mul_window##NLV(c, &R, &P, k, sb);                                         \
            else if ((MODE) == 1) pt_mul_base##NLV(c, &R, k, sb);                                          \
            else pt_mul_wnaf##NLV(c, &R, &P, k, sb);                                                       \
            fe##NLV x, y;                                                                                  \
            if (!pt_to_affine##NLV(c, &x, &y, &R)) { if (j->o2) j->o2[i] = 1; continue; }                  \
            f_to_be##NLV(&c->fp, out, &x, fb);                                                             \
            f_to_be##NLV(&c->fp, out + fb, &y, fb);                                                        \
        }                                                                                                  \
    }
/* mode 0 = &Point * &Scalar (fixed window), 1 = mul_base (comb), 2 = mul_vartime (wNAF) */
static void j_wei(job* j, size_t lo, size_t hi) {
    int mode = j->mode;
    switch (j->curve) {
        case 0: WEI_BODY(4, &P256, mode) break;
        case 1: WEI_BODY(6, &P384, mode) break;
        case 2: WEI_BODY(6, &BLSG1, mode) break;
        case 3: WEI_BODY(4, &K256, mode) break;
    }
}
/* PointAffine::decompress over a batch: a = x (fb bytes BE), b = sign bytes, o = x || y, o2 = present */
#define DECOMP_BODY(NLV, CURVE)                                                                            \
    {                                                                                                      \
        const curve##NLV* c = (CURVE);                                                                     \
        int fb = c->fbytes;                                                                                \
        for (size_t i = lo; i < hi; i++) {                                                                 \
            u8* out = j->o + (size_t)2 * fb * i;                                                           \
            memset(out, 0, (size_t)2 * fb);                                                                \
            j->o2[i] = 0;                                                                                  \
            fe##NLV x, y;                                                                                  \
            if (!f_from_be##NLV(&c->fp, &x, j->a + (size_t)fb * i, fb)) continue;                          \
            if (!pt_decompress##NLV(c, &y, &x, j->b[i] != 0)) continue;                                    \
            f_to_be##NLV(&c->fp, out, &x, fb);                                                             \
            f_to_be##NLV(&c->fp, out + fb, &y, fb);                                                        \
            j->o2[i] = 1;                                                                                  \
        }                                                                                                  \
    }
static void j_decompress(job* j, size_t lo, size_t hi) {
    switch (j->curve) {
        case 0: DECOMP_BODY(4, &P256) break;
        case 1: DECOMP_BODY(6, &P384) break;
        case 2: DECOMP_BODY(6, &BLSG1) break;
        case 3: DECOMP_BODY(4, &K256) break;
    }
}

/* ---- BLS12-381 G1 wire format and subgroup test -------------------------------------------
 * bls12_381/serialize.rs: flags in the top three bits of byte 0 (0x80 compressed, 0x40 infinity,
 * 0x20 y is the larger root: y > (p-1)/2); g1.rs:55-108: sigma(x, y) = (beta x, y),
 * mul_by_abs_x = double-and-add over |x| = 0xd201000000010000, P in G1 <=> sigma(P) = -[x^2]P. */
static fe6 BLS_BETA;          /* set in do_init from the cube roots of unity: the one with sigma(G) = [-x^2]G */
static u64 BLS_HALF_P[6];     /* (p - 1) / 2 */
#define BLS_ABS_X 0xd201000000010000ULL
static void bls_mul_by_abs_x(pt6* r, const pt6* p) {
    pt6 acc = *p;
    for (int i = 62; i >= 0; i--) {
        pt_dbl6(&BLSG1, &acc, &acc);
        if ((BLS_ABS_X >> i) & 1) pt_add6(&BLSG1, &acc, &acc, p);
    }
    *r = acc;
}
static int bls_pt_equiv(const pt6* a, const pt6* b) { /* projective equality (projective.rs:133) */
    const field6* F = &BLSG1.fp;
    fe6 l, r;
    f_mul6(F, &l, &a->X, &b->Z); f_mul6(F, &r, &b->X, &a->Z);
    if (!f_eq6(&l, &r)) return 0;
    f_mul6(F, &l, &a->Y, &b->Z); f_mul6(F, &r, &b->Y, &a->Z);
    return f_eq6(&l, &r);
}
static int bls_in_subgroup(const pt6* p) {
    pt6 t, u, s = *p;
    bls_mul_by_abs_x(&t, p);
    bls_mul_by_abs_x(&u, &t);
    f_neg6(&BLSG1.fp, &u.Y, &u.Y);
    f_mul6(&BLSG1.fp, &s.X, &s.X, &BLS_BETA);
    return bls_pt_equiv(&s, &u);
}
/* beta and (p-1)/2 are derived, not embedded: beta = g^((p-1)/3) != 1 for the first small g, then the
 * one of {beta, beta^2} that makes the generator pass the test (g1.rs:440-450 pins it the same way);
 * tests/test_oracle_golden.py compares it with the reference's BETA_BYTES */
static void bls_extra_init(void) {
    const field6* F = &BLSG1.fp;
    u64 e[6], pm1[6], one[6] = {1, 0, 0, 0, 0, 0};
    sub_raw6(pm1, F->p, one);
    for (int i = 0; i < 6; i++) BLS_HALF_P[i] = (pm1[i] >> 1) | (i + 1 < 6 ? pm1[i + 1] << 63 : 0);
    unsigned __int128 rem = 0;
    for (int i = 5; i >= 0; i--) {
        unsigned __int128 cur = (rem << 64) | pm1[i];
        e[i] = (u64)(cur / 3);
        rem = cur % 3;
    }
    for (u64 g = 2; g < 32; g++) {
        fe6 base, acc = F->r1;
        memset(&base, 0, sizeof base);
        base.v[0] = g;
        f_mul6(F, &base, &base, &F->r2);
        for (int i = 6 * 64 - 1; i >= 0; i--) {
            f_sqr6(F, &acc, &acc);
            if ((e[i / 64] >> (i % 64)) & 1) f_mul6(F, &acc, &acc, &base);
        }
        if (f_eq6(&acc, &F->r1)) continue;
        BLS_BETA = acc;
        if (!bls_in_subgroup(&BLSG1.G)) f_sqr6(F, &BLS_BETA, &acc);
        break;
    }
}
void orc_bls12_381_beta(u8* out48) {
    orc_init();
    f_to_be6(&BLSG1.fp, out48, &BLS_BETA, 48);
}
static int bls_y_is_largest(const fe6* y) {
    u64 w[6];
    f_to_raw6(&BLSG1.fp, w, y);
    return !ge_raw6(BLS_HALF_P, w);   /* w > (p-1)/2 */
}
/* a = 48-byte encodings, mode = check_subgroup; o = x || y (96 B), o2 = Some/None */
static void j_bls_from_compressed(job* j, size_t lo, size_t hi) {
    const curve6* c = &BLSG1;
    for (size_t i = lo; i < hi; i++) {
        const u8* enc = j->a + 48 * i;
        u8* out = j->o + 96 * i;
        memset(out, 0, 96);
        j->o2[i] = 0;
        u8 flags = enc[0] & 0xE0;
        if (!(flags & 0x80) || (flags & 0x40)) continue;   /* not compressed / the identity: None */
        u8 buf[48];
        memcpy(buf, enc, 48);
        buf[0] &= 0x1F;
        fe6 x, y;
        if (!f_from_be6(&c->fp, &x, buf, 48)) continue;
        if (!pt_decompress6(c, &y, &x, 0)) continue;
        if (bls_y_is_largest(&y) != ((flags & 0x20) ? 1 : 0)) f_neg6(&c->fp, &y, &y);
        if (j->mode) {
            pt6 P;
            P.X = x; P.Y = y; P.Z = c->fp.r1;
            if (!bls_in_subgroup(&P)) continue;
        }
        f_to_be6(&c->fp, out, &x, 48);
        f_to_be6(&c->fp, out + 48, &y, 48);
        j->o2[i] = 1;
    }
}
/* a = x || y (96 B), c = infinity flags (may be null); o = 48-byte encodings (Point::to_compressed) */
static void j_bls_to_compressed(job* j, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
        u8* out = j->o + 48 * i;
        if (j->c && j->c[i]) { memset(out, 0, 48); out[0] = 0xC0; continue; }
        memcpy(out, j->a + 96 * i, 48);
        u64 w[6];
        raw_from_be6(w, j->a + 96 * i + 48, 48);
        out[0] |= 0x80 | (ge_raw6(BLS_HALF_P, w) ? 0 : 0x20);
    }
}

static void j_ecdsa(job* j, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
        int bad = 0;
        if (j->curve == 0) j->o[i] = (u8)ecdsa_verify4(&P256, j->a + 64 * i, j->b + 32 * i, j->c + 64 * i, &bad);
        else j->o[i] = (u8)ecdsa_verify6(&P384, j->a + 96 * i, j->b + 48 * i, j->c + 96 * i, &bad);
        if (bad) BAD(j, i, ORC_BAD_POINT);
    }
}

/* ---- exports (ctypes): every function returns -1, or the index of the first invalid element and
 * its reason in *code (1 = non-canonical scalar, 2 = bad point) ---------------------------------- */
long orc_ed25519_mul_base(const u8* k_le, size_t n, u8* xy_le, int nthreads, int* code) {
    job t = {0};
    t.fn = j_ed_mul_base; t.a = k_le; t.o = xy_le;
    return run_jobs(&t, n, nthreads, code);
}
long orc_ed25519_mul(const u8* k_le, const u8* xy_in, size_t n, u8* xy_out, int nthreads, int* code) {
    job t = {0};
    t.fn = j_ed_mul; t.a = k_le; t.b = xy_in; t.o = xy_out;
    return run_jobs(&t, n, nthreads, code);
}
long orc_ed25519_verify_prehashed(const u8* a_enc, const u8* r_enc, const u8* s_le, const u8* k_le, size_t n, u8* ok,
                                  int nthreads) {
    job t = {0};
    t.fn = j_ed_verify; t.a = a_enc; t.b = r_enc; t.c = s_le; t.d = k_le; t.o = ok;
    return run_jobs(&t, n, nthreads, 0);
}
long orc_x25519(const u8* k, const u8* u, size_t n, u8* out, int nthreads) {
    job t = {0};
    t.fn = j_x25519; t.a = k; t.b = u; t.o = out;
    return run_jobs(&t, n, nthreads, 0);
}
long orc_x448(const u8* k, const u8* u, size_t n, u8* out, int nthreads) {
    job t = {0};
    t.fn = j_x448; t.a = k; t.b = u; t.o = out;
    return run_jobs(&t, n, nthreads, 0);
}
/* mode: 0 fixed-window (Point * Scalar), 1 comb (mul_base; xy ignored), 2 wNAF (mul_vartime) */
long orc_wei_mul(int curve, int mode, const u8* k_be, const u8* xy_be, const u8* inf_in, size_t n, u8* out_xy, u8* out_inf,
                 int nthreads, int* code) {
    if (curve < 0 || curve > 3 || mode < 0 || mode > 2) return -2;
    job t = {0};
    t.fn = j_wei; t.curve = curve; t.mode = mode; t.a = k_be; t.b = xy_be; t.c = inf_in; t.o = out_xy; t.o2 = out_inf;
    return run_jobs(&t, n, nthreads, code);
}
long orc_ecdsa_verify_hashed(int curve, const u8* q_xy, const u8* z_be, const u8* rs_be, size_t n, u8* ok, int nthreads,
                             int* code) {
    if (curve < 0 || curve > 1) return -2;
    job t = {0};
    t.fn = j_ecdsa; t.curve = curve; t.a = q_xy; t.b = z_be; t.c = rs_be; t.o = ok;
    return run_jobs(&t, n, nthreads, code);
}
/* comb-table entry (window i, digit j) as affine bytes, for pinning against params/comb/<curve>.rs */
/* PointAffine::decompress batch: sign[i] = 0 Positive (even y) / 1 Negative (odd y); ok[i] = present */
long orc_wei_decompress(int curve, const u8* x_be, const u8* sign, size_t n, u8* out_xy, u8* ok, int nthreads) {
    job j = {0};
    j.fn = j_decompress; j.curve = curve; j.a = x_be; j.b = sign; j.o = out_xy; j.o2 = ok;
    return run_jobs(&j, n, nthreads, 0);
}
long orc_bls12_381_g1_from_compressed(const u8* enc, size_t n, int check_subgroup, u8* out_xy, u8* ok, int nthreads) {
    job j = {0};
    j.fn = j_bls_from_compressed; j.mode = check_subgroup; j.a = enc; j.o = out_xy; j.o2 = ok;
    return run_jobs(&j, n, nthreads, 0);
}
long orc_bls12_381_g1_to_compressed(const u8* xy, const u8* inf, size_t n, u8* enc, int nthreads) {
    job j = {0};
    j.fn = j_bls_to_compressed; j.a = xy; j.c = inf; j.o = enc;
    return run_jobs(&j, n, nthreads, 0);
}
void orc_ed25519_comb_entry(int i, int j, u8* xy_le) {
    orc_init();
    ge_to_affine_bytes(xy_le, &ED_COMB[i][j]);
}
void orc_wei_comb_entry(int curve, int i, int j, u8* xy_be) {
    orc_init();
    if (curve == 0 || curve == 3) {
        const curve4* c4 = curve == 0 ? &P256 : &K256;
        f_to_be4(&c4->fp, xy_be, &c4->comb[i * 16 + j].X, 32);
        f_to_be4(&c4->fp, xy_be + 32, &c4->comb[i * 16 + j].Y, 32);
    } else {
        const curve6* c = curve == 1 ? &P384 : &BLSG1;
        f_to_be6(&c->fp, xy_be, &c->comb[i * 16 + j].X, 48);
        f_to_be6(&c->fp, xy_be + 48, &c->comb[i * 16 + j].Y, 48);
    }
}
