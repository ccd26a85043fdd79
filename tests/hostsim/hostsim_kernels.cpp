// Host simulation of the batch kernel bodies (tests only; never linked into libeccbatch.so).
// Runs the per-thread bodies of eccoxide_b200/csrc/kernels.cuh in a plain loop with the
// emulated carry flag, reproducing the launch structure of eccbatch.cu (kernel A -> planes ->
// batch inversion -> wire bytes).
#define ECB_HOSTSIM 1
#include "../../eccoxide_b200/csrc/kernels.cuh"
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
    for (size_t e = 0; e < ntab; e++) ed25519_table_point_body(e, ntab, W, nwin, planes.data());
    size_t T = inv_threads(ntab);
    FinEdNiels fin{planes.data(), ntab, table};
    for (size_t t = 0; t < T; t++) batch_inv_body<F25519>(t, T, ntab, planes.data(), pf.data(), fin);
}

unsigned long long hs_ed25519_mul_base(const u32* k, size_t n, int W, const u32* table, u32* out, int compressed) {
    int nwin = (254 + W - 1) / W;
    std::vector<u32> planes(3 * 8 * n), pf(8 * n);
    unsigned long long st = ~0ull;
    for (size_t i = 0; i < n; i++) ed25519_mul_base_body(i, n, k, table, W, nwin, planes.data(), &st);
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
    std::vector<u32> planes(3 * N * n), pf(N * n), tbl(8 * 3 * N);
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
    }
    return 0;
}
