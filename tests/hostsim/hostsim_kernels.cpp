// Host simulation of the batch kernel bodies (tests only; never linked into libeccbatch.so).
// Runs the per-thread bodies of eccoxide_b200/csrc/kernels.cuh in a plain loop with the
// emulated carry flag, reproducing the launch structure of eccbatch.cu (kernel A -> planes ->
// batch inversion -> wire bytes).
#define ECB_HOSTSIM 1
#include "../../eccoxide_b200/csrc/kernels.cuh"
#include "../../eccoxide_b200/csrc/ct.cuh"
#include "../../eccoxide_b200/csrc/ristretto.cuh"
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }   // one thread at a time here
#include "../../eccoxide_b200/csrc/msm.cuh"
#include <vector>
#include <string.h>
using namespace ecb;

static size_t inv_threads(size_t n) { size_t T = (n + 3) / 4; return T ? T : 1; }

extern "C" {

int hs_ed25519_table_entries(int W) { int nwin = (254 + W - 1) / W; return nwin << (W - 1); }

void hs_ed25519_build_table(int W, u32* table) {
    int nwin = (254 + W - 1) / W;
    size_t ntab = (size_t)nwin << (W - 1);
    std::vector<u32> planes(3 * 8 * ntab), pf(8 * ntab);
    std::vector<u32> bases((size_t)nwin * 32);
    for (int i = 0; i < nwin; i++) ed25519_window_base_body(i, W, bases.data());
    for (size_t e = 0; e < ntab; e++) ed25519_table_point_body(e, ntab, W, nwin, bases.data(), planes.data());
    size_t T = inv_threads(ntab);
    FinEdNiels fin{planes.data(), ntab, table};
    for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, ntab, planes.data(), pf.data(), fin);
}

unsigned long long hs_ed25519_mul_base(const u32* k, size_t n, int W, const u32* table, u32* out, int compressed) {
    int nwin = (254 + W - 1) / W;
    std::vector<u32> planes(3 * 8 * n), pf(8 * n);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) ed25519_mul_base_body<false>(i, n, k, table, W, nwin, 24, planes.data(), &st);
    size_t T = inv_threads(n);
    if (compressed) {
        FinEdCompressed fin{planes.data(), n, out};
        for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, n, planes.data(), pf.data(), fin);
    } else {
        FinEdXY fin{planes.data(), n, out};
        for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, n, planes.data(), pf.data(), fin);
    }
    return st;
}

// constant-time form (ct.cuh): masked scans of the W = 4 comb, Fermat inversion in the batch inversion
unsigned long long hs_ed25519_mul_base_ct(const u32* k, size_t n, const u32* table_w4, u32* out, int compressed) {
    std::vector<u32> planes(3 * 8 * n), pf(8 * n);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) ed25519_mul_base_ct_body<false>(i, n, k, table_w4, planes.data(), &st);
    size_t T = inv_threads(n);
    if (compressed) {
        FinEdCompressed fin{planes.data(), n, out};
        for (size_t t = 0; t < T; t++) batch_inv_body<F25519, FinEdCompressed, true>(t, T, n, planes.data(), pf.data(), fin);
    } else {
        FinEdXY fin{planes.data(), n, out};
        for (size_t t = 0; t < T; t++) batch_inv_body<F25519, FinEdXY, true>(t, T, n, planes.data(), pf.data(), fin);
    }
    return st;
}

// lane-split comb of the fused small-batch kernel (fused.cuh), without the warp: the partial sums of
// `lanes` lanes are added in the butterfly's order, then the usual affine finish
unsigned long long hs_ed25519_mul_base_lanes(const u32* k, size_t n, int W, const u32* table, int lanes, u32* out) {
    int nwin = (254 + W - 1) / W;
    std::vector<u32> planes(3 * 8 * n), pf(8 * n);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) {
        u32 kk[9];
        ed25519_load_scalar<false>(kk, i, k, &st);
        std::vector<ge_p3> part(lanes);
        for (int l = 0; l < lanes; l++) ed25519_comb_partial(part[l], kk, table, W, nwin, 24, l, lanes, lanes > 1);
        for (int d = 1; d < lanes; d <<= 1) {
            std::vector<ge_p3> nxt(lanes);
            for (int l = 0; l < lanes; l++) ge_add_p3<true>(nxt[l], part[l], part[l ^ d]);
            part = nxt;
        }
        plane_st<8>(planes.data() + 0 * 8 * n, n, i, part[0].X.v);
        plane_st<8>(planes.data() + 1 * 8 * n, n, i, part[0].Y.v);
        plane_st<8>(planes.data() + 2 * 8 * n, n, i, part[0].Z.v);
    }
    size_t T = inv_threads(n);
    FinEdXY fin{planes.data(), n, out};
    for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, n, planes.data(), pf.data(), fin);
    return st;
}

void hs_x25519_base(const u32* k, size_t n, int W, const u32* table, u32* out) {
    int nwin = (254 + W - 1) / W;
    std::vector<u32> planes(3 * 8 * n), pf(8 * n);
    for (size_t i = 0; i < n; i++) x25519_base_body(i, n, k, table, W, nwin, 24, planes.data());
    size_t T = inv_threads(n);
    FinEdMontU fin{planes.data(), n, out};
    for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, n, planes.data(), pf.data(), fin);
}

unsigned long long hs_ed25519_mul(const u32* k, const u32* pts, size_t n, u32* out) {
    std::vector<u32> planes(3 * 8 * n), pf(8 * n), tbl(8 * 32);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) ed25519_mul_body(i, n, k, pts, tbl.data(), planes.data(), &st);
    size_t T = inv_threads(n);
    FinEdXY fin{planes.data(), n, out};
    for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, n, planes.data(), pf.data(), fin);
    return st;
}

void hs_x25519(const u32* k, const u32* u, size_t n, u32* out) {
    std::vector<u32> planes(3 * 8 * n), pf(8 * n);
    for (size_t i = 0; i < n; i++) x25519_body(i, n, k, u, planes.data());
    size_t T = inv_threads(n);
    FinX25519 fin{planes.data(), n, out};
    for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, n, planes.data(), pf.data(), fin);
}
}

template <class C>
static unsigned long long wei_mul_run(const u32* k, const u32* pts, const unsigned char* inf_in, size_t n, u32* out,
                                      unsigned char* inf) {
    constexpr int N = C::F::N;
    std::vector<u32> planes(3 * N * n), pf(N * n), tbl(WeiJ<C>::TBL * 5 * N);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) wei_mul_body<C>(i, n, k, pts, inf_in, tbl.data(), planes.data(), &st);
    size_t T = inv_threads(n);
    FinWeiXY<C> fin{planes.data(), n, out, inf};
    for (size_t t = 0; t < T; t++) batch_inv_body<typename C::F>(t, T, n, planes.data(), pf.data(), fin);
    return st;
}

extern "C" unsigned long long hs_wei_mul(int curve, const u32* k, const u32* pts, const unsigned char* inf_in, size_t n,
                                         u32* out, unsigned char* inf) {
    switch (curve) {
        case 0: return wei_mul_run<CurveP256>(k, pts, inf_in, n, out, inf);
        case 1: return wei_mul_run<CurveP384>(k, pts, inf_in, n, out, inf);
        case 2: return wei_mul_run<CurveBLSG1>(k, pts, inf_in, n, out, inf);
        case 3: return wei_mul_run<CurveK256>(k, pts, inf_in, n, out, inf);
    }
    return 0;
}

// Jacobian doubling on arbitrary (loose) coordinate words: which = 0: dbl() (every fold checked); 1: dbl_z<false> (the same);
// 2: dbl_z<true>, the lazy form, returning its carry / borrow counters (nonzero = the result must be discarded and redone)
template <class C>
static unsigned wei_dbl_run(const u32* in, u32* out, int which) {
    typedef WeiJ<C> J;
    constexpr int N = C::F::N;
    typename J::pt p, r;
    memcpy(p.X.v, in, 4 * N); memcpy(p.Y.v, in + N, 4 * N); memcpy(p.Z.v, in + 2 * N, 4 * N);
    unsigned flags = 0;
    if (which == 0) J::dbl(r, p);
    else if (which == 1) {
        typename C::F::lazy z;
        J::template dbl_z<false>(r, p, z);
        flags = z.c | z.b;   // the checked forms never count
    } else {
        typename C::F::lazy z;
        J::template dbl_impl<true>(r, p, z);
        flags = z.c | z.b;
    }
    memcpy(out, r.X.v, 4 * N); memcpy(out + N, r.Y.v, 4 * N); memcpy(out + 2 * N, r.Z.v, 4 * N);
    return flags;
}
extern "C" unsigned hs_wei_dbl(int curve, const u32* in, u32* out, int which) {
    switch (curve) {
        case 0: return wei_dbl_run<CurveP256>(in, out, which);
        case 1: return wei_dbl_run<CurveP384>(in, out, which);
        case 2: return wei_dbl_run<CurveBLSG1>(in, out, which);
        case 3: return wei_dbl_run<CurveK256>(in, out, which);
    }
    return 0;
}

// ---- kernels2: X448, ECDSA verify, Ed25519 verify ------------------------------------------
#include "../../eccoxide_b200/csrc/kernels2.cuh"
extern "C" {
void hs_fe448(int op, const u32* a, const u32* b, u32* r) {
    fe448 x, y, z;
    memcpy(x.v, a, 56); memcpy(y.v, b, 56);
    switch (op) {
        case 0: F448::mul(z, x, y); break;
        case 1: F448::sqr(z, x); break;
        case 2: F448::add(z, x, y); break;
        case 3: F448::sub(z, x, y); break;
        case 5: F448::freeze(z, x); break;
        case 6: F448::invert(z, x); break;
        case 7: F448::mul_small(z, x, b[0]); break;
    }
    memcpy(r, z.v, 56);
}
void hs_x448(const u32* k, const u32* u, size_t n, u32* out) {
    std::vector<u32> planes(3 * 14 * n), pf(14 * n);
    for (size_t i = 0; i < n; i++) x448_body(i, n, k, u, planes.data());
    size_t T = inv_threads(n);
    FinX448 fin{planes.data(), n, out};
    for (size_t t = 0; t < T; t++) batch_inv_body<F448>(t, T, n, planes.data(), pf.data(), fin);
}
void hs_sha512(const unsigned char* msg, size_t len, unsigned char* digest) {
    sha512_bytes(digest, len, [&](size_t pos) -> unsigned char { return msg[pos]; });
}
void hs_ecdsa_hash_z(const unsigned char* msgs, const unsigned long long* off, size_t n, int hash, int SB, unsigned char* z) {
    for (size_t i = 0; i < n; i++) ecdsa_hash_z_body(i, msgs, off, hash, SB, z);
}
void hs_ed25519_hash_k(const unsigned char* a, const unsigned char* sig, const unsigned char* msgs, const unsigned long long* off,
                       size_t n, u32* r, u32* s, u32* k) {
    for (size_t i = 0; i < n; i++) ed25519_hash_k_body(i, a, sig, msgs, off, r, s, k);
}
void hs_ed25519_verify(const u32* a, const u32* r, const u32* s, const u32* k, size_t n, int W, const u32* table,
                       unsigned char* ok) {
    int nwin = (254 + W - 1) / W;
    std::vector<u32> tbl(8 * 32), planes(3 * 8 * n), pf(8 * n);
    for (size_t i = 0; i < n; i++) ed25519_verify_body(i, n, a, s, k, table, W, nwin, 24, tbl.data(), planes.data(), ok);
    size_t T = inv_threads(n);
    FinEdVerify fin{planes.data(), n, r, ok};
    for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, n, planes.data(), pf.data(), fin);
}
}
template <class C>
static std::vector<u32> wei_build_comb(int W, int nwin) {
    constexpr int N = C::F::N;
    size_t ntab = (size_t)nwin << (W - 1);
    std::vector<u32> planes(3 * N * ntab), pf(N * ntab), table(2 * N * ntab);
    std::vector<u32> bases((size_t)nwin * 5 * N);
    for (int i = 0; i < nwin; i++) wei_window_base_body<C>(i, W, bases.data());
    for (size_t e = 0; e < ntab; e++) wei_table_point_body<C>(e, ntab, W, nwin, bases.data(), planes.data());
    size_t T = inv_threads(ntab);
    FinWeiTable<C> fin{planes.data(), ntab, table.data()};
    for (size_t t = 0; t < T; t++) batch_inv_body<typename C::F>(t, T, ntab, planes.data(), pf.data(), fin);
    return table;
}
template <class C>
static unsigned long long wei_mul_base_run(const u32* k, size_t n, int W, u32* out, unsigned char* inf) {
    constexpr int N = C::F::N;
    int nwin = (C::SBITS + 1 + W - 1) / W;
    std::vector<u32> table = wei_build_comb<C>(W, nwin);
    std::vector<u32> planes(3 * N * n), pf(N * n);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) wei_mul_base_body<C>(i, n, k, table.data(), W, nwin, planes.data(), &st);
    size_t T = inv_threads(n);
    FinWeiXY<C> fin{planes.data(), n, out, inf};
    for (size_t t = 0; t < T; t++) batch_inv_body<typename C::F>(t, T, n, planes.data(), pf.data(), fin);
    return st;
}
// constant-time form (ct.cuh): W = 4 comb scanned with masks, complete projective additions, Fermat inversion
template <class C>
static unsigned long long wei_mul_base_ct_run(const u32* k, size_t n, u32* out, unsigned char* inf) {
    constexpr int N = C::F::N;
    std::vector<u32> table = wei_build_comb<C>(ECB_CT_W, WeiCt<C>::NWIN);
    std::vector<u32> planes(3 * N * n), pf(N * n);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) wei_mul_base_ct_body<C>(i, n, k, table.data(), planes.data(), &st);
    size_t T = inv_threads(n);
    FinWeiXY<C> fin{planes.data(), n, out, inf};
    for (size_t t = 0; t < T; t++) batch_inv_body<typename C::F, FinWeiXY<C>, true>(t, T, n, planes.data(), pf.data(), fin);
    return st;
}
extern "C" unsigned long long hs_wei_mul_base_ct(int curve, const u32* k, size_t n, u32* out, unsigned char* inf) {
    if (curve == 0) return wei_mul_base_ct_run<CurveP256>(k, n, out, inf);
    return wei_mul_base_ct_run<CurveP384>(k, n, out, inf);
}
extern "C" unsigned long long hs_wei_mul_base(int curve, const u32* k, size_t n, int W, u32* out, unsigned char* inf) {
    switch (curve) {
        case 0: return wei_mul_base_run<CurveP256>(k, n, W, out, inf);
        case 1: return wei_mul_base_run<CurveP384>(k, n, W, out, inf);
        case 2: return wei_mul_base_run<CurveBLSG1>(k, n, W, out, inf);
        case 3: return wei_mul_base_run<CurveK256>(k, n, W, out, inf);
    }
    return 0;
}
template <class C>
static unsigned long long ecdsa_run(const u32* q, const u32* z, const u32* rs, size_t n, unsigned char* ok) {
    constexpr int N = C::F::N, NS = C::FN::N;
    std::vector<u32> planes(3 * N * n), pf((N > NS ? N : NS) * n), aux(3 * NS * n), tbl(WeiJ<C>::TBL * 5 * N);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) ecdsa_prep_body<C>(i, n, z, rs, aux.data(), ok);
    size_t T = inv_threads(n);
    FinScalarInv<C> f1{aux.data(), n};
    for (size_t t = 0; t < T; t++) batch_inv_body<typename C::FN>(t, T, n, aux.data(), pf.data(), f1);
    const int W = 5, nwin = (C::SBITS + 1 + W - 1) / W;
    std::vector<u32> gtab = wei_build_comb<C>(W, nwin);
    for (size_t i = 0; i < n; i++) ecdsa_main_body<C>(i, n, q, z, rs, aux.data(), ok, tbl.data(), gtab.data(), W, nwin, planes.data(), &st);
    FinEcdsa<C> f2{planes.data(), n, rs, ok};
    for (size_t t = 0; t < T; t++) batch_inv_body<typename C::F>(t, T, n, planes.data(), pf.data(), f2);
    return st;
}
extern "C" unsigned long long hs_ecdsa_verify(int curve, const u32* q, const u32* z, const u32* rs, size_t n, unsigned char* ok) {
    if (curve == 0) return ecdsa_run<CurveP256>(q, z, rs, n, ok);
    return ecdsa_run<CurveP384>(q, z, rs, n, ok);
}

// ---- kernels3: PointAffine::decompress, BLS12-381 G1 standard encodings ---------------------
#include "../../eccoxide_b200/csrc/kernels3.cuh"

// BLS12-381 G1 through the endomorphism (option bls12_381_g1_glv): the split and the two-scalar window loop
extern "C" void hs_bls_glv_split(const u32* k, u32* k1, u32* k2) { bls_glv_split(k1, k2, k); }
extern "C" unsigned long long hs_bls_g1_mul_glv(const u32* k, const u32* pts, const unsigned char* inf_in, size_t n, u32* out, unsigned char* inf) {
    typedef CurveBLSG1 C;
    constexpr int N = C::F::N;
    std::vector<u32> planes(3 * N * n), pf(N * n), tbl(WeiJ<C>::TBL * 5 * N);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) wei_mul_glv_body<C>(i, n, k, pts, inf_in, tbl.data(), planes.data(), &st);
    size_t T = inv_threads(n);
    FinWeiXY<C> fin{planes.data(), n, out, inf};
    for (size_t t = 0; t < T; t++) batch_inv_body<typename C::F>(t, T, n, planes.data(), pf.data(), fin);
    return st;
}
extern "C" void hs_wei_decompress(int curve, const u32* x, const unsigned char* sign, size_t n, u32* out, unsigned char* ok) {
    for (size_t i = 0; i < n; i++) {
        switch (curve) {
            case 0: wei_decompress_body<CurveP256>(i, x, sign, out, ok); break;
            case 1: wei_decompress_body<CurveP384>(i, x, sign, out, ok); break;
            case 2: wei_decompress_body<CurveBLSG1>(i, x, sign, out, ok); break;
            case 3: wei_decompress_body<CurveK256>(i, x, sign, out, ok); break;
        }
    }
}
extern "C" void hs_bls_g1_from_compressed(const u32* enc, size_t n, int check, u32* out, unsigned char* ok) {
    for (size_t i = 0; i < n; i++) bls_g1_from_compressed_body(i, enc, check, out, ok);
}
extern "C" void hs_bls_g1_to_compressed(const u32* xy, const unsigned char* inf, size_t n, u32* enc) {
    for (size_t i = 0; i < n; i++) bls_g1_to_compressed_body(i, xy, inf, enc);
}

// ---- Ed25519 key generation and signing (kernels2.cuh) ---------------------------------------
extern "C" void hs_ed25519_public_from_seed(const unsigned char* seeds, size_t n, int W, const u32* table, u32* pub) {
    int nwin = (254 + W - 1) / W;
    std::vector<u32> a(8 * n), prefix(8 * n), planes(3 * 8 * n), pf(8 * n);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) ed25519_expand_body(i, seeds, a.data(), prefix.data());
    for (size_t i = 0; i < n; i++) ed25519_mul_base_body<false>(i, n, a.data(), table, W, nwin, 24, planes.data(), &st);
    size_t T = inv_threads(n);
    FinEdCompressed fin{planes.data(), n, pub};
    for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, n, planes.data(), pf.data(), fin);
}
extern "C" void hs_ed25519_sign(const unsigned char* seeds, const unsigned char* pub, const unsigned char* msgs,
                                const unsigned long long* off, size_t n, int W, const u32* table, unsigned char* sig) {
    int nwin = (254 + W - 1) / W;
    std::vector<u32> a(8 * n), prefix(8 * n), r(8 * n), planes(3 * 8 * n), pf(8 * n);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) ed25519_expand_body(i, seeds, a.data(), prefix.data());
    for (size_t i = 0; i < n; i++) ed25519_sign_nonce_body(i, prefix.data(), msgs, off, r.data());
    for (size_t i = 0; i < n; i++) ed25519_mul_base_body<false>(i, n, r.data(), table, W, nwin, 24, planes.data(), &st);
    size_t T = inv_threads(n);
    FinEdCompressed fin{planes.data(), n, (u32*)sig, 16};
    for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, n, planes.data(), pf.data(), fin);
    for (size_t i = 0; i < n; i++) ed25519_sign_finish_body(i, sig, pub, msgs, off, a.data(), r.data());
}

// ---- ECDSA sign_hashed (kernels2.cuh) -------------------------------------------------------
template <class C>
static void ecdsa_sign_run(const u32* d, const u32* k, const u32* z, size_t n, u32* rs, unsigned char* ok) {
    constexpr int N = C::F::N;
    constexpr int NS = C::FN::N;
    std::vector<u32> sp(3 * NS * n), pf((N > NS ? N : NS) * n), kg(2 * N * n);
    std::vector<unsigned char> kinf(n);
    for (size_t i = 0; i < n; i++) ecdsa_sign_prep_body<C>(i, n, d, k, z, sp.data(), ok);
    size_t T = inv_threads(n);
    FinScalarInv<C> fin1{sp.data(), n};
    for (size_t t = 0; t < T; t++) batch_inv_body<typename C::FN>(t, T, n, sp.data(), pf.data(), fin1);
    wei_mul_base_run<C>(k, n, 5, kg.data(), kinf.data());
    for (size_t i = 0; i < n; i++) ecdsa_sign_finish_body<C>(i, n, d, z, kg.data(), kinf.data(), sp.data(), rs, ok);
}
extern "C" void hs_ecdsa_sign(int curve, const u32* d, const u32* k, const u32* z, size_t n, u32* rs, unsigned char* ok) {
    if (curve == 0) ecdsa_sign_run<CurveP256>(d, k, z, n, rs, ok);
    else ecdsa_sign_run<CurveP384>(d, k, z, n, rs, ok);
}

// ristretto255 encodings (ristretto.cuh)
extern "C" void hs_ristretto255_decompress(const u32* enc, size_t n, u32* xy, unsigned char* ok) {
    for (size_t i = 0; i < n; i++) ristretto255_decompress_body(i, enc, xy, ok);
}
extern "C" void hs_ristretto255_compress(const u32* xy, size_t n, u32* enc) {
    for (size_t i = 0; i < n; i++) ristretto255_compress_body(i, xy, enc);
}

// multi-scalar multiplication (msm.cuh): the kernels' bodies in launch order, window width c, segment length S and
// running-sum chunk CH given; the warp-level sum of a heavy bucket's pieces (k_msm_finish in tu_wei.inc) is replayed lane by lane
template <class C>
static unsigned long long msm_run(const u32* k, const u32* pts, size_t n, int c, size_t S, u32 CHmax, u32 heavy, u32* out, unsigned char* inf) {
    constexpr int N = C::F::N;
    typedef Msm<C> M;
    typedef typename M::J J;
    const int nwin = (C::SBITS + 1 + c - 1) / c;
    const u32 NB = 1u << (c - 1);
    const size_t nb = (size_t)nwin * NB;
    const size_t nseg = ((size_t)nwin * n + S - 1) / S;
    const u32 CH = NB < CHmax ? NB : CHmax, NC = NB / CH;
    const size_t nc = (size_t)nwin * NC;
    std::vector<u32> pm(n * 2 * N), dig((size_t)nwin * n), idx((size_t)nwin * n), hist(nb, 0), offs(nb, 0), cursor(nb, 0),
        bsum(nb * M::PW, 0xdeadbeefu), csum(nc * M::PW), pieces(2 * nseg * M::PW, 0xdeadbeefu), part(M::PW);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) msm_prepare_body<C>(i, n, k, pts, c, nwin, NB, pm.data(), dig.data(), hist.data(), &st);
    u32 run = 0;
    for (size_t b = 0; b < nb; b++) { offs[b] = run; run += hist[b]; }
    for (size_t i = 0; i < n; i++) msm_scatter_body(i, n, nwin, NB, dig.data(), offs.data(), cursor.data(), idx.data());
    for (size_t t = 0; t < nseg; t++) msm_segment_body<C>(t, S, nb, offs.data(), hist.data(), idx.data(), pm.data(), bsum.data(), pieces.data());
    for (size_t b = 0; b < nb; b++) {
        typename J::pt mid;
        bool has_mid = false;
        if (hist[b]) {
            const size_t bs = offs[b], be = bs + hist[b], tf = bs / S, tl = (be - 1) / S;
            if (tl - tf > heavy) {
                typename J::pt lane[32];
                for (int l = 0; l < 32; l++) {
                    J::set_inf(lane[l]);
                    msm_pieces<C>(lane[l], pieces.data(), tf + 1 + l, tl + 1, 32);
                }
                for (int d = 16; d >= 1; d >>= 1)
                    for (int l = 0; l < d; l++) M::add_pts(lane[l], lane[l], lane[l + d]);
                mid = lane[0];
                has_mid = true;
            }
        }
        msm_finish_body<C>(b, S, offs.data(), hist.data(), pieces.data(), bsum.data(), has_mid ? &mid : nullptr);
    }
    for (size_t t = 0; t < nc; t++) msm_chunk_body<C>(t, NB, CH, bsum.data(), csum.data());
    for (u32 half = NC >> 1; half >= 1; half >>= 1)
        for (size_t t = 0; t < (size_t)nwin * half; t++) msm_reduce_body<C>(t, NC, half, csum.data());
    msm_window_body<C>(c, nwin, NC, csum.data(), part.data());
    std::vector<u32> planes(3 * N), pf(N);
    msm_combine_body<C>(1, part.data(), planes.data());
    FinWeiXY<C> fin{planes.data(), 1, out, inf};
    batch_inv_body<typename C::F>(0, 1, 1, planes.data(), pf.data(), fin);
    return st;
}
extern "C" unsigned long long hs_wei_msm2(int curve, const u32* k, const u32* pts, size_t n, int c, size_t S, u32 CH, u32 heavy, u32* out, unsigned char* inf) {
    if (curve == 2) return msm_run<CurveBLSG1>(k, pts, n, c, S, CH, heavy, out, inf);
    return msm_run<CurveK256>(k, pts, n, c, S, CH, heavy, out, inf);
}
extern "C" unsigned long long hs_wei_msm(int curve, const u32* k, const u32* pts, size_t n, int c, u32* out, unsigned char* inf) {
    if (curve == 2) return msm_run<CurveBLSG1>(k, pts, n, c, 8, 32, 16, out, inf);
    return msm_run<CurveK256>(k, pts, n, c, 8, 32, 16, out, inf);
}
