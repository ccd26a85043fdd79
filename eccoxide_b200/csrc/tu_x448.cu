// tu_x448.cu — X448 ladder kernel.
#include "tu_common.cuh"
#include "dev_ops.h"

static __global__ void __launch_bounds__(ECB_TPB, 3) k_x448(size_t n, const u32* scalars, const u32* us, u32* planes) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) x448_body(idx, n, scalars, us, planes);
}

int dev_x448(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_u, size_t n, u32* d_out, cudaStream_t s) {
    TRY(ensure(ctx, d.cur->planes, n * 3 * 14 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 14 * sizeof(u32)));
    TRY(reset_status(ctx, d, s));  // these kernels report no per-element errors; keep the word clean for ecb_dev_status
    u32* planes = (u32*)d.cur->planes.p;
    prof_mark(ctx, d, s, 0);
    k_x448<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_k, d_u, planes);
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);
    FinX448 fin{planes, n, d_out};
    int rc = launch_batch_inv<F448, FinX448>(ctx, d, n, planes, (u32*)d.cur->pf.p, fin, s);
    prof_mark(ctx, d, s, 2);
    return rc;
}
