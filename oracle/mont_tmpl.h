/* mont_tmpl.h — word-by-word Montgomery field on NL 64-bit limbs + short-Weierstrass projective
 * arithmetic over it.  Included once per limb count (NL = 4, 6) by ecc_oracle.c.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see ecc_oracle.c header).
 *
 * Restates, in plain C with unsigned __int128:
 *   - the fiat word-by-word Montgomery backends src/curve/fiat/p256_64.rs:265 (mul), :859 (add),
 *     :941 (sub), :1013 (opp), :1084/:1259 (from/to_montgomery), :1643/:1757 (to/from_bytes) and
 *     their p384 / bls12_381 / *_scalar twins: R = 2^(64 NL), values canonical in [0, p);
 *   - src/curve/projective.rs: add_different_am3 :340, double_am3 :586, add_different_a0 :268,
 *     double_a0 :544 (Renes-Costello-Batina complete formulas), select_from_table :427,
 *     scalar_mul_fixed_window_{am3,a0} :871/:842, mul_base_table_{am3,a0} :965/:945, wnaf :65,
 *     scalar_mul_wnaf_{am3,a0} :771/:745, to_affine_ct :675.
 */
#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define T(name) CAT(name, NL)

typedef struct { u64 v[NL]; } T(fe);
typedef struct {
    u64 p[NL];
    u64 ninv; /* -p^-1 mod 2^64 */
    T(fe) r1, r2;
    int bits;
} T(field);

static int T(ge_raw)(const u64* a, const u64* b) { /* a >= b */
    for (int i = NL - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static u64 T(add_raw)(u64* r, const u64* a, const u64* b) {
    u128 c = 0;
    for (int i = 0; i < NL; i++) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static u64 T(sub_raw)(u64* r, const u64* a, const u64* b) {
    u64 bw = 0;
    for (int i = 0; i < NL; i++) {
        u128 d = (u128)a[i] - b[i] - bw;
        r[i] = (u64)d;
        bw = (u64)(d >> 64) & 1;
    }
    return bw;
}
static void T(f_add)(const T(field)* f, T(fe)* r, const T(fe)* a, const T(fe)* b) {
    u64 t[NL], d[NL];
    u64 c = T(add_raw)(t, a->v, b->v);
    u64 bw = T(sub_raw)(d, t, f->p);
    const u64* s = (c | (bw ^ 1)) ? d : t;
    for (int i = 0; i < NL; i++) r->v[i] = s[i];
}
static void T(f_sub)(const T(field)* f, T(fe)* r, const T(fe)* a, const T(fe)* b) {
    u64 t[NL];
    u64 bw = T(sub_raw)(t, a->v, b->v);
    if (bw) T(add_raw)(t, t, f->p);
    for (int i = 0; i < NL; i++) r->v[i] = t[i];
}
static void T(f_neg)(const T(field)* f, T(fe)* r, const T(fe)* a) {
    T(fe) z;
    memset(&z, 0, sizeof z);
    T(f_sub)(f, r, &z, a);
}
/* CIOS Montgomery product a*b*R^-1 mod p (fiat_*_mul) */
static void T(f_mul)(const T(field)* f, T(fe)* r, const T(fe)* a, const T(fe)* b) {
    u64 t[NL + 2];
    for (int i = 0; i < NL + 2; i++) t[i] = 0;
#pragma GCC unroll 8
    for (int i = 0; i < NL; i++) {
        u128 c = 0;
#pragma GCC unroll 8
        for (int j = 0; j < NL; j++) {
            c += (u128)a->v[j] * b->v[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[NL];
        t[NL] = (u64)c;
        t[NL + 1] = (u64)(c >> 64);
        u64 m = t[0] * f->ninv;
        c = (u128)m * f->p[0] + t[0];
        c >>= 64;
#pragma GCC unroll 8
        for (int j = 1; j < NL; j++) {
            c += (u128)m * f->p[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[NL];
        t[NL - 1] = (u64)c;
        t[NL] = t[NL + 1] + (u64)(c >> 64);
    }
    u64 d[NL];
    u64 bw = T(sub_raw)(d, t, f->p);
    const u64* s = (t[NL] | (bw ^ 1)) ? d : t;
    for (int i = 0; i < NL; i++) r->v[i] = s[i];
}
static void T(f_sqr)(const T(field)* f, T(fe)* r, const T(fe)* a) { T(f_mul)(f, r, a, a); } /* fiat square costs the same as mul */
static int T(f_is_zero)(const T(fe)* a) {
    u64 o = 0;
    for (int i = 0; i < NL; i++) o |= a->v[i];
    return o == 0;
}
static int T(f_eq)(const T(fe)* a, const T(fe)* b) {
    u64 o = 0;
    for (int i = 0; i < NL; i++) o |= a->v[i] ^ b->v[i];
    return o == 0;
}
static void T(f_select)(T(fe)* r, u64 take_a, const T(fe)* a, const T(fe)* b) { /* ct_select */
    u64 m = 0 - (take_a & 1);
    for (int i = 0; i < NL; i++) r->v[i] = (a->v[i] & m) | (b->v[i] & ~m);
}
/* bytes (big-endian, len = 8*NL or shorter) -> raw limbs; returns 0 if >= p when `check` */
static void T(raw_from_be)(u64* w, const u8* b, int len) {
    for (int i = 0; i < NL; i++) w[i] = 0;
    for (int i = 0; i < len; i++) w[(len - 1 - i) / 8] |= (u64)b[i] << (8 * ((len - 1 - i) % 8));
}
static void T(raw_to_be)(u8* b, const u64* w, int len) {
    for (int i = 0; i < len; i++) b[i] = (u8)(w[(len - 1 - i) / 8] >> (8 * ((len - 1 - i) % 8)));
}
/* FieldElement::from_bytes (field_macros.rs:604-660): canonical check then to_montgomery */
static int T(f_from_be)(const T(field)* f, T(fe)* r, const u8* b, int len) {
    T(fe) t;
    T(raw_from_be)(t.v, b, len);
    if (T(ge_raw)(t.v, f->p)) return 0;
    T(f_mul)(f, r, &t, &f->r2);
    return 1;
}
/* FieldElement::to_bytes (field_macros.rs:661-676): from_montgomery then serialise */
static void T(f_to_raw)(const T(field)* f, u64* w, const T(fe)* a) {
    T(fe) one, t;
    memset(&one, 0, sizeof one);
    one.v[0] = 1;
    T(f_mul)(f, &t, a, &one);
    for (int i = 0; i < NL; i++) w[i] = t.v[i];
}
static void T(f_to_be)(const T(field)* f, u8* b, const T(fe)* a, int len) {
    u64 w[NL];
    T(f_to_raw)(f, w, a);
    T(raw_to_be)(b, w, len);
}
/* a^(p-2) with a 4-bit fixed window over the exponent (the reference uses per-prime addition
 * chains of the same size class: p256r1.rs:49-65 255 S + 12 M, p384r1.rs:50-69; bls fp.rs:55 uses
 * safegcd).  0 -> 0.  Any exponentiation to p-2 yields the same field element. */
static void T(f_inv)(const T(field)* f, T(fe)* r, const T(fe)* a) {
    T(fe) tab[16], acc;
    u64 e[NL];
    u64 two[NL];
    for (int i = 0; i < NL; i++) two[i] = 0;
    two[0] = 2;
    T(sub_raw)(e, f->p, two);
    tab[0] = f->r1;
    tab[1] = *a;
    for (int i = 2; i < 16; i++) T(f_mul)(f, &tab[i], &tab[i - 1], a);
    acc = f->r1;
    for (int i = NL * 16 - 1; i >= 0; i--) {
        for (int s = 0; s < 4; s++) T(f_sqr)(f, &acc, &acc);
        unsigned d = (unsigned)(e[i / 16] >> (4 * (i % 16))) & 15u;
        if (d) T(f_mul)(f, &acc, &acc, &tab[d]);
    }
    *r = acc;
}
static void T(field_init)(T(field)* f, const u8* p_be, int len) {
    T(raw_from_be)(f->p, p_be, len);
    u64 x = 1; /* Newton: x = p^-1 mod 2^64 */
    for (int i = 0; i < 6; i++) x *= 2 - f->p[0] * x;
    f->ninv = 0 - x;
    f->bits = 0;
    for (int i = NL * 64 - 1; i >= 0; i--)
        if ((f->p[i / 64] >> (i % 64)) & 1) { f->bits = i + 1; break; }
    /* r1 = 2^(64 NL) mod p, r2 = 2^(128 NL) mod p by repeated doubling of 1 */
    T(fe) t;
    memset(&t, 0, sizeof t);
    t.v[0] = 1;
    for (int i = 0; i < 128 * NL; i++) {
        T(f_add)(f, &t, &t, &t);
        if (i == 64 * NL - 1) f->r1 = t;
    }
    f->r2 = t;
}

/* ------------------------------------------------------------------------------------------
 * homogeneous projective points (X : Y : Z), INFINITY = (0 : 1 : 0)   (projective.rs:28, :152)
 * ---------------------------------------------------------------------------------------- */
typedef struct { T(fe) X, Y, Z; } T(pt);
typedef struct {
    T(field) fp, fn;
    T(fe) b, b3;
    T(pt) G;
    int a0;            /* 1: a = 0 (BLS12-381 G1), 0: a = -3 */
    int fbytes, sbytes;
    int nwin;          /* comb windows = 2 * sbytes (params/comb: COMB_WINDOWS) */
    T(pt)* comb;       /* [nwin][16], index 0 = INFINITY (projective.rs:451 build_comb_table) */
} T(curve);

static void T(pt_inf)(const T(curve)* c, T(pt)* r) {
    memset(r, 0, sizeof *r);
    r->Y = c->fp.r1;
}
static void T(pt_select)(T(pt)* r, u64 take_a, const T(pt)* a, const T(pt)* b) {
    T(f_select)(&r->X, take_a, &a->X, &b->X);
    T(f_select)(&r->Y, take_a, &a->Y, &b->Y);
    T(f_select)(&r->Z, take_a, &a->Z, &b->Z);
}
#define FM(r, a, b) T(f_mul)(F, &(r), &(a), &(b))
#define FA(r, a, b) T(f_add)(F, &(r), &(a), &(b))
#define FS(r, a, b) T(f_sub)(F, &(r), &(a), &(b))
/* RCB Algorithm 4 (a = -3): 12 M + 2 m_b, projective.rs:340-423 */
static void T(pt_add_am3)(const T(curve)* c, T(pt)* r, const T(pt)* p, const T(pt)* q) {
    const T(field)* F = &c->fp;
    T(fe) t0, t1, t2, t3, t4, X3, Y3, Z3;
    FM(t0, p->X, q->X); FM(t1, p->Y, q->Y); FM(t2, p->Z, q->Z);
    FA(t3, p->X, p->Y); FA(t4, q->X, q->Y); FM(t3, t3, t4);
    FA(t4, t0, t1); FS(t3, t3, t4); FA(t4, p->Y, p->Z);
    FA(X3, q->Y, q->Z); FM(t4, t4, X3); FA(X3, t1, t2);
    FS(t4, t4, X3); FA(X3, p->X, p->Z); FA(Y3, q->X, q->Z);
    FM(X3, X3, Y3); FA(Y3, t0, t2); FS(Y3, X3, Y3);
    FM(Z3, c->b, t2); FS(X3, Y3, Z3); FA(Z3, X3, X3);
    FA(X3, X3, Z3); FS(Z3, t1, X3); FA(X3, t1, X3);
    FM(Y3, c->b, Y3); FA(t1, t2, t2); FA(t2, t1, t2);
    FS(Y3, Y3, t2); FS(Y3, Y3, t0); FA(t1, Y3, Y3);
    FA(Y3, t1, Y3); FA(t1, t0, t0); FA(t0, t1, t0);
    FS(t0, t0, t2); FM(t1, t4, Y3); FM(t2, t0, Y3);
    FM(Y3, X3, Z3); FA(Y3, Y3, t2); FM(X3, t3, X3);
    FS(X3, X3, t1); FM(Z3, t4, Z3); FM(t1, t3, t0);
    FA(Z3, Z3, t1);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
/* RCB Algorithm 6 (a = -3): 8 M + 3 S + 2 m_b, projective.rs:586-646 */
static void T(pt_dbl_am3)(const T(curve)* c, T(pt)* r, const T(pt)* p) {
    const T(field)* F = &c->fp;
    T(fe) t0, t1, t2, t3, X3, Y3, Z3;
    FM(t0, p->X, p->X); FM(t1, p->Y, p->Y); FM(t2, p->Z, p->Z);
    FM(t3, p->X, p->Y); FA(t3, t3, t3); FM(Z3, p->X, p->Z);
    FA(Z3, Z3, Z3); FM(Y3, c->b, t2); FS(Y3, Y3, Z3);
    FA(X3, Y3, Y3); FA(Y3, X3, Y3); FS(X3, t1, Y3);
    FA(Y3, t1, Y3); FM(Y3, X3, Y3); FM(X3, X3, t3);
    FA(t3, t2, t2); FA(t2, t2, t3); FM(Z3, c->b, Z3);
    FS(Z3, Z3, t2); FS(Z3, Z3, t0); FA(t3, Z3, Z3);
    FA(Z3, Z3, t3); FA(t3, t0, t0); FA(t0, t3, t0);
    FS(t0, t0, t2); FM(t0, t0, Z3); FA(Y3, Y3, t0);
    FM(t0, p->Y, p->Z); FA(t0, t0, t0); FM(Z3, t0, Z3);
    FS(X3, X3, Z3); FM(Z3, t0, t1); FA(Z3, Z3, Z3);
    FA(Z3, Z3, Z3);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
/* RCB Algorithm 7 (a = 0): 12 M + 2 m_3b, projective.rs:268-338 */
static void T(pt_add_a0)(const T(curve)* c, T(pt)* r, const T(pt)* p, const T(pt)* q) {
    const T(field)* F = &c->fp;
    T(fe) t0, t1, t2, t3, t4, X3, Y3, Z3;
    FM(t0, p->X, q->X); FM(t1, p->Y, q->Y); FM(t2, p->Z, q->Z);
    FA(t3, p->X, p->Y); FA(t4, q->X, q->Y); FM(t3, t3, t4);
    FA(t4, t0, t1); FS(t3, t3, t4); FA(t4, p->Y, p->Z);
    FA(X3, q->Y, q->Z); FM(t4, t4, X3); FA(X3, t1, t2);
    FS(t4, t4, X3); FA(X3, p->X, p->Z); FA(Y3, q->X, q->Z);
    FM(X3, X3, Y3); FA(Y3, t0, t2); FS(Y3, X3, Y3);
    FA(X3, t0, t0); FA(t0, X3, t0); FM(t2, c->b3, t2);
    FA(Z3, t1, t2); FS(t1, t1, t2); FM(Y3, c->b3, Y3);
    FM(X3, t4, Y3); FM(t2, t3, t1); FS(X3, t2, X3);
    FM(Y3, Y3, t0); FM(t1, t1, Z3); FA(Y3, t1, Y3);
    FM(t0, t0, t3); FM(Z3, Z3, t4); FA(Z3, Z3, t0);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
/* RCB Algorithm 9 (a = 0): 6 M + 2 S + 1 m_3b, projective.rs:544-583 */
static void T(pt_dbl_a0)(const T(curve)* c, T(pt)* r, const T(pt)* p) {
    const T(field)* F = &c->fp;
    T(fe) t0, t1, t2, X3, Y3, Z3;
    FM(t0, p->Y, p->Y); FA(Z3, t0, t0); FA(Z3, Z3, Z3);
    FA(Z3, Z3, Z3); FM(t1, p->Y, p->Z); FM(t2, p->Z, p->Z);
    FM(t2, c->b3, t2); FM(X3, t2, Z3); FA(Y3, t0, t2);
    FM(Z3, t1, Z3); FA(t1, t2, t2); FA(t2, t1, t2);
    FS(t0, t0, t2); FM(Y3, t0, Y3); FA(Y3, X3, Y3);
    FM(t1, p->X, p->Y); FM(X3, t0, t1); FA(X3, X3, X3);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}
#undef FM
#undef FA
#undef FS
static void T(pt_add)(const T(curve)* c, T(pt)* r, const T(pt)* p, const T(pt)* q) {
    if (c->a0) T(pt_add_a0)(c, r, p, q); else T(pt_add_am3)(c, r, p, q);
}
static void T(pt_dbl)(const T(curve)* c, T(pt)* r, const T(pt)* p) {
    if (c->a0) T(pt_dbl_a0)(c, r, p); else T(pt_dbl_am3)(c, r, p);
}
/* select_from_table (projective.rs:427): full 16-entry scan */
static void T(pt_lookup16)(const T(curve)* c, T(pt)* r, const T(pt)* table, unsigned idx) {
    T(pt) acc;
    T(pt_inf)(c, &acc);
    for (unsigned j = 0; j < 16; j++) T(pt_select)(&acc, j == idx, &table[j], &acc);
    *r = acc;
}
/* scalar_mul_fixed_window_{am3,a0} (projective.rs:871 / :842); n = big-endian scalar bytes */
static void T(pt_mul_window)(const T(curve)* c, T(pt)* r, const T(pt)* p, const u8* n, int nlen) {
    T(pt) table[16], q, sel;
    T(pt_inf)(c, &table[0]);
    table[1] = *p;
    T(pt_dbl)(c, &table[2], p);
    for (int d = 3; d < 16; d++) T(pt_add)(c, &table[d], &table[d - 1], p);
    T(pt_inf)(c, &q);
    for (int i = 0; i < nlen; i++) {
        unsigned idx[2] = {(unsigned)n[i] >> 4, (unsigned)n[i] & 15u};
        for (int h = 0; h < 2; h++) {
            for (int s = 0; s < 4; s++) T(pt_dbl)(c, &q, &q);
            T(pt_lookup16)(c, &sel, table, idx[h]);
            T(pt_add)(c, &q, &q, &sel);
        }
    }
    *r = q;
}
/* mul_base_table_{am3,a0} (projective.rs:965 / :945) */
static void T(pt_mul_base)(const T(curve)* c, T(pt)* r, const u8* n, int nlen) {
    T(pt) q, sel;
    T(pt_inf)(c, &q);
    for (int i = 0; i < c->nwin; i++) {
        u8 byte = n[nlen - 1 - i / 2];
        unsigned digit = (i % 2 == 0) ? (byte & 15u) : (byte >> 4);
        T(pt_lookup16)(c, &sel, c->comb + (size_t)i * 16, digit);
        T(pt_add)(c, &q, &q, &sel);
    }
    *r = q;
}
/* wnaf (projective.rs:65-109) over big-endian bytes; returns the digit count */
static int T(wnaf)(signed char* naf, const u8* n, int nlen, int w) {
    u8 k[8 * NL + 2];
    for (int i = 0; i < nlen; i++) k[i] = n[nlen - 1 - i];
    k[nlen] = 0;
    int klen = nlen + 1, cnt = 0;
    int width = 1 << w, half = 1 << (w - 1);
    for (;;) {
        int any = 0;
        for (int i = 0; i < klen; i++) any |= k[i];
        if (!any) break;
        int digit = 0;
        if (k[0] & 1) {
            int m = k[0] & (width - 1);
            digit = m >= half ? m - width : m;
            int carry = -digit;
            for (int i = 0; carry != 0 && i < klen; i++) {
                int v = (int)k[i] + carry;
                k[i] = (u8)(v & 0xff);
                carry = v >> 8;
            }
        }
        naf[cnt++] = (signed char)digit;
        u8 prev = 0;
        for (int i = klen - 1; i >= 0; i--) {
            u8 cur = k[i];
            k[i] = (u8)((cur >> 1) | (prev << 7));
            prev = cur & 1;
        }
    }
    return cnt;
}
/* scalar_mul_wnaf_{am3,a0} (projective.rs:771 / :745), w = 5 (WNAF_W) */
static void T(pt_mul_wnaf)(const T(curve)* c, T(pt)* r, const T(pt)* p, const u8* n, int nlen) {
    signed char naf[64 * NL + 16];
    int cnt = T(wnaf)(naf, n, nlen, 5);
    T(pt) dbl, table[8], q, neg;
    T(pt_dbl)(c, &dbl, p);
    table[0] = *p;
    for (int i = 1; i < 8; i++) T(pt_add)(c, &table[i], &table[i - 1], &dbl);
    T(pt_inf)(c, &q);
    for (int i = cnt - 1; i >= 0; i--) {
        T(pt_dbl)(c, &q, &q);
        int d = naf[i];
        if (d > 0) {
            T(pt_add)(c, &q, &q, &table[d >> 1]);
        } else if (d < 0) {
            neg = table[(-d) >> 1];
            T(f_neg)(&c->fp, &neg.Y, &neg.Y);
            T(pt_add)(c, &q, &q, &neg);
        }
    }
    *r = q;
}
/* to_affine_ct (projective.rs:675): returns 0 for the identity */
static int T(pt_to_affine)(const T(curve)* c, T(fe)* x, T(fe)* y, const T(pt)* p) {
    if (T(f_is_zero)(&p->Z)) return 0;
    T(fe) zi;
    T(f_inv)(&c->fp, &zi, &p->Z);
    T(f_mul)(&c->fp, x, &p->X, &zi);
    if (y) T(f_mul)(&c->fp, y, &p->Y, &zi);
    return 1;
}
/* affine::Point::from_coordinate (affine.rs:77-104): canonical coordinates on the curve */
static int T(pt_from_xy_be)(const T(curve)* c, T(pt)* r, const u8* xy) {
    const T(field)* F = &c->fp;
    if (!T(f_from_be)(F, &r->X, xy, c->fbytes)) return 0;
    if (!T(f_from_be)(F, &r->Y, xy + c->fbytes, c->fbytes)) return 0;
    r->Z = F->r1;
    T(fe) l, rr, t;
    T(f_mul)(F, &l, &r->Y, &r->Y);
    T(f_mul)(F, &rr, &r->X, &r->X);
    T(f_mul)(F, &rr, &rr, &r->X);
    if (!c->a0) {
        T(f_add)(F, &t, &r->X, &r->X);
        T(f_add)(F, &t, &t, &r->X);
        T(f_sub)(F, &rr, &rr, &t);
    }
    T(f_add)(F, &rr, &rr, &c->b);
    return T(f_eq)(&l, &rr);
}
/* FieldElement::sqrt (p256r1.rs:68, p384r1.rs:71, bls12_381/fp.rs:64): candidate = a^((p+1)/4)
 * (all three primes are 3 mod 4; the reference walks per-prime addition chains / `power`, the
 * value is the same), present iff candidate^2 == a. */
static int T(f_sqrt)(const T(field)* f, T(fe)* r, const T(fe)* a) {
    u64 e[NL], one[NL];
    for (int i = 0; i < NL; i++) one[i] = 0;
    one[0] = 1;
    T(add_raw)(e, f->p, one);                  /* p + 1 < 2^(64 NL) for these primes */
    for (int i = 0; i < NL; i++) e[i] = (e[i] >> 2) | (i + 1 < NL ? e[i + 1] << 62 : 0);
    T(fe) acc = f->r1, chk;
    for (int i = NL * 64 - 1; i >= 0; i--) {
        T(f_sqr)(f, &acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) T(f_mul)(f, &acc, &acc, a);
    }
    T(f_sqr)(f, &chk, &acc);
    *r = acc;
    return T(f_eq)(&chk, a);
}
/* FieldElement::sign (field_macros.rs:557): bit 0 of the canonical value */
static int T(f_odd)(const T(field)* f, const T(fe)* a) {
    u64 w[NL];
    T(f_to_raw)(f, w, a);
    return (int)(w[0] & 1);
}
/* affine::Point::decompress (affine.rs:48-60): y^2 = x^3 + a x + b, root of the requested parity */
static int T(pt_decompress)(const T(curve)* c, T(fe)* y, const T(fe)* x, int want_odd) {
    const T(field)* F = &c->fp;
    T(fe) rr, t;
    T(f_mul)(F, &rr, x, x);
    T(f_mul)(F, &rr, &rr, x);
    if (!c->a0) {
        T(f_add)(F, &t, x, x);
        T(f_add)(F, &t, &t, x);
        T(f_sub)(F, &rr, &rr, &t);
    }
    T(f_add)(F, &rr, &rr, &c->b);
    int present = T(f_sqrt)(F, y, &rr);
    if (T(f_odd)(F, y) != (want_odd ? 1 : 0)) T(f_neg)(F, y, y);
    return present;
}
static void T(curve_init)(T(curve)* c, const u8* p_be, const u8* n_be, const u8* b_be, const u8* gx_be, const u8* gy_be,
                          int fbytes, int sbytes, int a0) {
    memset(c, 0, sizeof *c);
    c->fbytes = fbytes; c->sbytes = sbytes; c->a0 = a0;
    T(field_init)(&c->fp, p_be, fbytes);
    {   /* the scalar modulus may be shorter than 8*NL bytes (BLS Fr on 4 limbs is handled by NL=4) */
        T(field_init)(&c->fn, n_be, sbytes);
    }
    T(f_from_be)(&c->fp, &c->b, b_be, fbytes);
    T(f_add)(&c->fp, &c->b3, &c->b, &c->b);
    T(f_add)(&c->fp, &c->b3, &c->b3, &c->b);
    T(f_from_be)(&c->fp, &c->G.X, gx_be, fbytes);
    T(f_from_be)(&c->fp, &c->G.Y, gy_be, fbytes);
    c->G.Z = c->fp.r1;
    /* comb table: entry [i][j] = j * 16^i * G, affine (Z = 1), [i][0] = INFINITY
     * (params/comb/<curve>.rs layout, sage/comb.sage:73-111; recomputed here, not embedded) */
    c->nwin = 2 * sbytes;
    c->comb = (T(pt)*)malloc(sizeof(T(pt)) * 16 * (size_t)c->nwin);
    T(pt) base = c->G;
    for (int i = 0; i < c->nwin; i++) {
        T(pt)* w = c->comb + (size_t)i * 16;
        T(pt_inf)(c, &w[0]);
        w[1] = base;
        for (int j = 2; j < 16; j++) T(pt_add)(c, &w[j], &w[j - 1], &base);
        for (int j = 1; j < 16; j++) {
            T(fe) x = c->fp.r1, y = c->fp.r1;   /* multiples of G below the group order are finite */
            T(pt_to_affine)(c, &x, &y, &w[j]);
            w[j].X = x; w[j].Y = y; w[j].Z = c->fp.r1;
        }
        T(pt) nb;
        T(pt_add)(c, &nb, &w[15], &base); /* 16 * base */
        T(fe) x = c->fp.r1, y = c->fp.r1;
        T(pt_to_affine)(c, &x, &y, &nb);
        base.X = x; base.Y = y; base.Z = c->fp.r1;
    }
}
#undef T
#undef CAT
#undef CAT_
