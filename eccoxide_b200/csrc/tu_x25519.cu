// tu_x25519.cu — X25519 ladder kernel.
#include "tu_common.cuh"
#include "dev_ops.h"

static __global__ void __launch_bounds__(ECB_TPB) k_x25519(size_t n, const u32* scalars, const u32* us, u32* planes) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) x25519_body(idx, n, scalars, us, planes);
}

int dev_x25519(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_u, size_t n, u32* d_out, cudaStream_t s) {
    TRY(ensure(ctx, d.cur->planes, n * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 8 * sizeof(u32)));
    TRY(reset_status(ctx, d, s));  // these kernels report no per-element errors; keep the word clean for ecb_dev_status
    u32* planes = (u32*)d.cur->planes.p;
    prof_mark(ctx, d, s, 0);
    k_x25519<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_k, d_u, planes);
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);
    FinX25519 fin{planes, n, d_out};
    int rc = launch_batch_inv<F25519, FinX25519>(ctx, d, n, planes, (u32*)d.cur->pf.p, fin, s);
    prof_mark(ctx, d, s, 2);
    return rc;
}
