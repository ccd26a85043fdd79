// tu_ed25519.cu — edwards25519 kernels: fixed-base comb (+ table builder), variable base, verify.
#include "tu_common.cuh"
#include "dev_ops.h"
#include "fused.cuh"
#include "ct.cuh"
#include "ristretto.cuh"

static __global__ void __launch_bounds__(ECB_TPB, 5) k_ed25519_mul_base(size_t n, const u32* scalars, const u32* table, int W,
                                                               int nwin, int stride, u32* planes, unsigned long long* status) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) ed25519_mul_base_body<false>(idx, n, scalars, table, W, nwin, stride, planes, status);
}
// small-batch form (fused.cuh): LANES lanes per scalar, affine conversion in the same launch.
// <FUSED_WIDE_T, 1>: one block per SM of up to 15 work warps + the inversion warp (512 threads, 128 registers:
// the register file is split over four schedulers, so 14 or 15 warps get no more registers than 16), the whole
// batch in one wave — 148 x 480 = 71040 >= 2^16 scalars, BASELINE configs[0]; <160, 3>: the same body with a
// large-batch launch shape (option ed25519_fused = 2, for measurement).
#define FUSED_WIDE_T 512
#define FUSED_WORK_T (FUSED_WIDE_T - 32)
template <int LANES, bool CLAMP, class FIN, int MAXT, int MINB>
static __global__ void __launch_bounds__(MAXT, MINB) k_ed25519_mul_base_fused(size_t n, const u32* scalars, const u32* table, int W, int nwin,
                                                                      int stride, FIN fin, unsigned long long* status,
                                                                      unsigned long long* trace) {
    __shared__ u32 sh[(FUSED_MAXW + 1) * 8];
    __shared__ u32 jump[SG_JUMP_WORDS];
    ed25519_mul_base_fused_block<LANES, CLAMP, FIN>(n, scalars, table, W, nwin, stride, fin, status, sh, jump, trace);
}
static __global__ void __launch_bounds__(ECB_TPB) k_ed25519_window_bases(int nwin, int W, u32* bases) {
    int i = (int)(blockIdx.x * ECB_TPB + threadIdx.x);
    if (i < nwin) ed25519_window_base_body(i, W, bases);
}
static __global__ void __launch_bounds__(ECB_TPB) k_ed25519_table_points(size_t ntab, int W, int nwin, const u32* bases, u32* planes) {
    size_t e = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (e < ntab) ed25519_table_point_body(e, ntab, W, nwin, bases, planes);
}
static __global__ void __launch_bounds__(ECB_TPB) k_ed25519_mul(size_t n, const u32* scalars, const u32* points, u32* scratch,
                                                          u32* planes, unsigned long long* status) {
    size_t t = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    u32* tbl = scratch + t * (8 * 32);
    for (size_t base = warp_next_chunk(status); base < n; base = warp_next_chunk(status)) {
        const size_t idx = base + (threadIdx.x & 31u);
        if (idx < n) ed25519_mul_body(idx, n, scalars, points, tbl, planes, status);
    }
}
static __global__ void __launch_bounds__(ECB_TPB, 3) k_ed25519_verify(size_t n, const u32* a_enc, const u32* s_le, const u32* k_le,
                                                             const u32* table, int W, int nwin, int stride, u32* scratch, u32* planes,
                                                             unsigned char* ok, unsigned long long* status) {
    size_t t = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    u32* tbl = scratch + t * (8 * 32);
    for (size_t base = warp_next_chunk(status); base < n; base = warp_next_chunk(status)) {
        const size_t idx = base + (threadIdx.x & 31u);
        if (idx < n) ed25519_verify_body(idx, n, a_enc, s_le, k_le, table, W, nwin, stride, tbl, planes, ok);
    }
}

// comb width actually used on this device: the option, or (option 0) the widest whose table and
// transient build buffers (~2.5x the table) fit comfortably in the free memory
static int ed_pick_w(ecb_ctx* ctx) {
    if (ctx->opt_ed_w) return (int)ctx->opt_ed_w;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 16;
    // W = 26 (10 windows, 32 GB table + 43 GB while it is built) is 7 % faster at 2^20 but is opt-in: ed25519_comb_w = 26
    if (free_b > ((size_t)40 << 30)) return 24;    // 11 windows, 8.9 GB table
    if (free_b > ((size_t)8 << 30)) return 20;
    return 16;
}
static int ed25519_build_table_w(ecb_ctx* ctx, DevCtx& d, int W);
static void ed_drop_table(DevCtx& d) {
    if (d.ed_table) cudaFree(d.ed_table);
    d.ed_table = nullptr;
    d.ed_w = d.ed_nwin = 0;
}
// In automatic mode (option 0) a width whose table or build buffers cannot be allocated falls back
// to the next narrower even width instead of failing the call.  On failure the context holds NO table
// (ed_table == nullptr, ed_w == ed_nwin == 0): the next call rebuilds or fails again, loudly.
int dev_ed25519_build_table(ecb_ctx* ctx, DevCtx& d, int W) {
    for (;;) {
        int rc = ed25519_build_table_w(ctx, d, W);
        if (rc == ECB_OK || ctx->opt_ed_w != 0 || W <= 16) return rc;
        (void)cudaGetLastError();
        for (DevBuf* b : {&d.cur->planes, &d.cur->pf}) {
            if (b->p) cudaFree(b->p);
            b->p = nullptr;
            b->cap = 0;
        }
        W -= 2;
    }
}
// Build a comb of width W into a fresh device buffer (entries `stride` words apart); the caller owns it.
static int ed25519_build_table_raw(ecb_ctx* ctx, DevCtx& d, int W, int stride, u32** out_table, int* out_nwin) {
    const int nwin = (254 + W - 1) / W;
    const size_t ntab = (size_t)nwin << (W - 1);
    u32* table = nullptr;
    u32* bases = nullptr;   // 2^(W i) * B per window, cached form (32 words each)
    auto build = [&]() -> int {
        CU(cudaMalloc(&table, ntab * (size_t)stride * sizeof(u32)));
        TRY(ensure(ctx, d.cur->planes, ntab * 3 * 8 * sizeof(u32)));
        TRY(ensure(ctx, d.cur->pf, ntab * 8 * sizeof(u32)));
        CU(cudaMalloc(&bases, (size_t)nwin * 32 * sizeof(u32)));
        if (stride != 24) CU(cudaMemsetAsync(table, 0, ntab * (size_t)stride * sizeof(u32), d.stream));
        k_ed25519_window_bases<<<grid_for((size_t)nwin), ECB_TPB, 0, d.stream>>>(nwin, W, bases);
        k_ed25519_table_points<<<grid_for(ntab), ECB_TPB, 0, d.stream>>>(ntab, W, nwin, bases, (u32*)d.cur->planes.p);
        ctx->launches += 2;
        CU(cudaGetLastError());
        FinEdNiels fin{(const u32*)d.cur->planes.p, ntab, table, (size_t)stride};
        TRY((launch_batch_inv<F25519, FinEdNiels>(ctx, d, ntab, (const u32*)d.cur->planes.p, (u32*)d.cur->pf.p, fin, d.stream)));
        CU(cudaStreamSynchronize(d.stream));
        return ECB_OK;
    };
    int rc = build();
    if (rc != ECB_OK) cudaStreamSynchronize(d.stream);   // nothing may still write into buffers about to be freed
    if (bases) cudaFree(bases);
    if (rc != ECB_OK) {
        if (table) cudaFree(table);
        return rc;
    }
    if (ntab * 96 > ((size_t)256 << 20)) {  // give the build buffers of a large table back
        for (DevBuf* b : {&d.cur->planes, &d.cur->pf}) {
            if (b->p) cudaFree(b->p);
            b->p = nullptr;
            b->cap = 0;
        }
    }
    *out_table = table;
    *out_nwin = nwin;
    return ECB_OK;
}
// The new table is published (table, width, window count, stride together) only when every step succeeded;
// any failure leaves the context without a table.
static int ed25519_build_table_w(ecb_ctx* ctx, DevCtx& d, int W) {
    const int stride = (int)ctx->opt_ed_stride;
    ed_drop_table(d);
    u32* table = nullptr;
    int nwin = 0;
    TRY(ed25519_build_table_raw(ctx, d, W, stride, &table, &nwin));
    d.ed_table = table;
    d.ed_w = W;
    d.ed_nwin = nwin;
    d.ed_stride = stride;
    return ECB_OK;
}
static inline bool ed_table_stale(ecb_ctx* ctx, const DevCtx& d) {
    return !d.ed_table || d.ed_nwin <= 0 || (ctx->opt_ed_w && d.ed_w != (int)ctx->opt_ed_w) || d.ed_stride != (int)ctx->opt_ed_stride;
}
// make sure the device holds a comb table of the configured shape (ecb_warm and the host entry points)
int dev_ed25519_table(ecb_ctx* ctx, DevCtx& d) {
    if (!ed_table_stale(ctx, d)) return ECB_OK;
    if (ctx->no_alloc) return set_err(ctx, ECB_ERR_NOT_READY, "Ed25519 comb table not built for the current options: call ecb_warm() first");
    return dev_ed25519_build_table(ctx, d, ed_pick_w(ctx));
}

// Launch shape of the fused small-batch kernel: LANES lanes per scalar and a block size such that the
// whole batch is one wave of one block (<= FUSED_WORK_T work threads) per SM.  tpb counts the work threads; the
// launch adds the inversion warp.  false: the batch is too large for it.
static bool fused_shape(const DevCtx& d, size_t n, long force_lanes, int& lanes, int& tpb, unsigned& grid) {
    const size_t cap = (size_t)d.sm_count * FUSED_WORK_T;
    if (force_lanes) {
        lanes = (int)force_lanes;
    } else {   // measured (profiles/r02_tune_ed25519.jsonl): extra lanes pay while they add warps to idle schedulers, i.e. up to ~4 warps per SM
        lanes = 8;
        while (lanes > 1 && n * lanes > (size_t)d.sm_count * 128) lanes >>= 1;
    }
    size_t threads = n * lanes;
    if (threads > cap) return false;
    size_t per_sm = (threads + d.sm_count - 1) / d.sm_count;
    tpb = (int)((per_sm + 31) / 32 * 32);
    if (tpb < 32) tpb = 32;
    grid = (unsigned)((threads + tpb - 1) / tpb);
    return true;
}
template <bool CLAMP, class FIN>
static int launch_fused(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, FIN fin, unsigned long long* status, cudaStream_t s, bool& done) {
    done = false;
    if (!ctx->opt_ed_fused || n == 0) return ECB_OK;
    int lanes, tpb;
    unsigned grid;
    const bool fits = fused_shape(d, n, ctx->opt_ed_lanes, lanes, tpb, grid);
    if (!fits && ctx->opt_ed_fused != 2) return ECB_OK;
    unsigned long long* trace = nullptr;
    if (ctx->opt_trace) {   // measurement only: 4 timestamps per block of the last fused launch (ecb_debug_fused_trace)
        size_t blocks = fits ? grid : grid_for(n);
        TRY(ensure(ctx, d.trace, blocks * 4 * sizeof(unsigned long long)));
        trace = (unsigned long long*)d.trace.p;
        d.trace_blocks = blocks;
    }
    prof_mark(ctx, d, s, 0);
    if (fits) {
        const int T = tpb + 32;   // + the inversion warp
        switch (lanes) {
            case 1: k_ed25519_mul_base_fused<1, CLAMP, FIN, FUSED_WIDE_T, 1><<<grid, T, 0, s>>>(n, d_k, d.ed_table, d.ed_w, d.ed_nwin, d.ed_stride, fin, status, trace); break;
            case 2: k_ed25519_mul_base_fused<2, CLAMP, FIN, FUSED_WIDE_T, 1><<<grid, T, 0, s>>>(n, d_k, d.ed_table, d.ed_w, d.ed_nwin, d.ed_stride, fin, status, trace); break;
            case 4: k_ed25519_mul_base_fused<4, CLAMP, FIN, FUSED_WIDE_T, 1><<<grid, T, 0, s>>>(n, d_k, d.ed_table, d.ed_w, d.ed_nwin, d.ed_stride, fin, status, trace); break;
            default: k_ed25519_mul_base_fused<8, CLAMP, FIN, FUSED_WIDE_T, 1><<<grid, T, 0, s>>>(n, d_k, d.ed_table, d.ed_w, d.ed_nwin, d.ed_stride, fin, status, trace); break;
        }
    } else {
        k_ed25519_mul_base_fused<1, CLAMP, FIN, ECB_TPB + 32, 3><<<grid_for(n), ECB_TPB + 32, 0, s>>>(n, d_k, d.ed_table, d.ed_w, d.ed_nwin, d.ed_stride, fin, status, trace);
    }
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);   // one kernel: the whole step is booked as "scalar_mult", the finisher share is 0
    prof_mark(ctx, d, s, 2);
    done = true;
    return ECB_OK;
}

int dev_ed25519_mul_base(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, bool compressed, cudaStream_t s,
                         size_t enc_stride_words) {
    TRY(dev_ed25519_table(ctx, d));
    TRY(reset_status(ctx, d, s));
    {
        bool done = false;
        if (compressed) {
            FusedEdCompressed fin{d_out, enc_stride_words ? enc_stride_words : 8};
            TRY((launch_fused<false, FusedEdCompressed>(ctx, d, d_k, n, fin, d.cur->d_status, s, done)));
        } else {
            FusedEdXY fin{d_out};
            TRY((launch_fused<false, FusedEdXY>(ctx, d, d_k, n, fin, d.cur->d_status, s, done)));
        }
        if (done) return ECB_OK;
    }
    TRY(ensure(ctx, d.cur->planes, n * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 8 * sizeof(u32)));
    u32* planes = (u32*)d.cur->planes.p;
    prof_mark(ctx, d, s, 0);
    k_ed25519_mul_base<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_k, d.ed_table, d.ed_w, d.ed_nwin, d.ed_stride, planes, d.cur->d_status);
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);
    int rc;
    if (compressed) {
        FinEdCompressed fin{planes, n, d_out, enc_stride_words ? enc_stride_words : 8};
        rc = launch_batch_inv<F25519, FinEdCompressed>(ctx, d, n, planes, (u32*)d.cur->pf.p, fin, s);
    } else {
        FinEdXY fin{planes, n, d_out};
        rc = launch_batch_inv<F25519, FinEdXY>(ctx, d, n, planes, (u32*)d.cur->pf.p, fin, s);
    }
    prof_mark(ctx, d, s, 2);
    return rc;
}


// ---- constant-time fixed base (ct.cuh): secret scalars ------------------------------------------------
// Persistent blocks: each stages the 48 KB W = 4 comb into shared memory once, then its threads walk the batch.
static __global__ void __launch_bounds__(ECB_TPB, 4) k_ed25519_mul_base_ct(size_t n, const u32* scalars, const u32* table, u32* planes,
                                                                       unsigned long long* status) {
    __shared__ u32 tbl[ECB_CT_ED_WORDS];
    for (int i = threadIdx.x; i < ECB_CT_ED_WORDS / 4; i += ECB_TPB)
        reinterpret_cast<uint4*>(tbl)[i] = reinterpret_cast<const uint4*>(table)[i];
    __syncthreads();
    // the order in which warps take their chunks depends on the progress of other warps, never on a scalar
    for (size_t base = warp_next_chunk(status); base < n; base = warp_next_chunk(status)) {
        const size_t idx = base + (threadIdx.x & 31u);
        if (idx < n) ed25519_mul_base_ct_body<false>(idx, n, scalars, tbl, planes, status);
    }
}
static int dev_ed25519_ct_table(ecb_ctx* ctx, DevCtx& d) {
    if (d.ed_ct_table) return ECB_OK;
    if (ctx->no_alloc) return set_err(ctx, ECB_ERR_NOT_READY, "constant-time Ed25519 comb not built: call ecb_warm() first");
    u32* table = nullptr;
    int nwin = 0;
    TRY(ed25519_build_table_raw(ctx, d, ECB_CT_W, 24, &table, &nwin));
    if (nwin != ECB_CT_ED_NWIN) {
        cudaFree(table);
        return set_err(ctx, ECB_ERR_CUDA, "constant-time comb: unexpected window count");
    }
    d.ed_ct_table = table;
    return ECB_OK;
}
// k * B for secret k: masked scans of a shared-memory comb, Fermat inversion.  ~9x the cost of the variable-time form.
int dev_ed25519_mul_base_ct(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, bool compressed, cudaStream_t s,
                            size_t enc_stride_words) {
    TRY(dev_ed25519_ct_table(ctx, d));
    TRY(reset_status(ctx, d, s));
    TRY(ensure(ctx, d.cur->planes, n * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 8 * sizeof(u32)));
    u32* planes = (u32*)d.cur->planes.p;
    unsigned g = persistent_grid(d, k_ed25519_mul_base_ct, n);
    prof_mark(ctx, d, s, 0);
    k_ed25519_mul_base_ct<<<g, ECB_TPB, 0, s>>>(n, d_k, d.ed_ct_table, planes, d.cur->d_status);
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);
    int rc;
    if (compressed) {
        FinEdCompressed fin{planes, n, d_out, enc_stride_words ? enc_stride_words : 8};
        rc = launch_batch_inv_ct<F25519, FinEdCompressed>(ctx, d, n, planes, (u32*)d.cur->pf.p, fin, s);
    } else {
        FinEdXY fin{planes, n, d_out};
        rc = launch_batch_inv_ct<F25519, FinEdXY>(ctx, d, n, planes, (u32*)d.cur->pf.p, fin, s);
    }
    prof_mark(ctx, d, s, 2);
    return rc;
}

static __global__ void __launch_bounds__(ECB_TPB, 5) k_x25519_base(size_t n, const u32* scalars, const u32* table, int W, int nwin,
                                                          int stride, u32* planes) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) x25519_base_body(idx, n, scalars, table, W, nwin, stride, planes);
}
int dev_x25519_base(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, cudaStream_t s) {
    TRY(dev_ed25519_table(ctx, d));
    if (d.ed_w * d.ed_nwin < 256) return set_err(ctx, ECB_ERR_INVALID_ARG, "x25519_base needs a comb covering 256 bits (ed25519_comb_w)");
    TRY(reset_status(ctx, d, s));
    {
        bool done = false;
        FusedEdMontU fin{d_out};
        TRY((launch_fused<true, FusedEdMontU>(ctx, d, d_k, n, fin, nullptr, s, done)));
        if (done) return ECB_OK;
    }
    TRY(ensure(ctx, d.cur->planes, n * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 8 * sizeof(u32)));
    u32* planes = (u32*)d.cur->planes.p;
    prof_mark(ctx, d, s, 0);
    k_x25519_base<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_k, d.ed_table, d.ed_w, d.ed_nwin, d.ed_stride, planes);
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);
    FinEdMontU fin{planes, n, d_out};
    int rc = launch_batch_inv<F25519, FinEdMontU>(ctx, d, n, planes, (u32*)d.cur->pf.p, fin, s);
    prof_mark(ctx, d, s, 2);
    return rc;
}

int dev_ed25519_mul(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_p, size_t n, u32* d_out, cudaStream_t s) {
    TRY(ensure(ctx, d.cur->planes, n * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 8 * sizeof(u32)));
    unsigned g = persistent_grid(d, k_ed25519_mul, n);
    TRY(ensure(ctx, d.cur->scratch, (size_t)g * ECB_TPB * 8 * 32 * sizeof(u32)));
    TRY(reset_status(ctx, d, s));
    u32* planes = (u32*)d.cur->planes.p;
    prof_mark(ctx, d, s, 0);
    k_ed25519_mul<<<g, ECB_TPB, 0, s>>>(n, d_k, d_p, (u32*)d.cur->scratch.p, planes, d.cur->d_status);
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);
    FinEdXY fin{planes, n, d_out};
    int rc = launch_batch_inv<F25519, FinEdXY>(ctx, d, n, planes, (u32*)d.cur->pf.p, fin, s);
    prof_mark(ctx, d, s, 2);
    return rc;
}

static __global__ void __launch_bounds__(ECB_TPB) k_ed25519_decompress(size_t n, const u32* enc, u32* out_xy, unsigned char* ok) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) ed25519_decompress_body(idx, enc, out_xy, ok);
}
int dev_ed25519_decompress(ecb_ctx* ctx, DevCtx& d, const u32* d_enc, size_t n, u32* d_out, unsigned char* d_ok, cudaStream_t s) {
    (void)d;
    k_ed25519_decompress<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_enc, d_out, d_ok);
    ctx->launches++;
    CU(cudaGetLastError());
    return ECB_OK;
}
static __global__ void __launch_bounds__(ECB_TPB) k_ed25519_hash_k(size_t n, const unsigned char* a_enc, const unsigned char* sig,
                                                             const unsigned char* msgs, const unsigned long long* off,
                                                             u32* r_out, u32* s_out, u32* k_out) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) ed25519_hash_k_body(idx, a_enc, sig, msgs, off, r_out, s_out, k_out);
}
// verification on raw messages: k = SHA-512(R || A || M) mod l on the device, then the prehashed path.
// d_msgs is addressed as d_msgs[off[i] .. off[i+1]) (the caller may pass a pointer already rebased).
int dev_ed25519_verify_msgs(ecb_ctx* ctx, DevCtx& d, const unsigned char* a, const unsigned char* sig, const unsigned char* d_msgs,
                            const unsigned long long* d_off, size_t n, unsigned char* ok, cudaStream_t s) {
    TRY(ensure(ctx, d.cur->aux, n * 3 * 32));
    u32* r = (u32*)d.cur->aux.p;
    u32* sl = r + n * 8;
    u32* kl = sl + n * 8;
    k_ed25519_hash_k<<<grid_for(n), ECB_TPB, 0, s>>>(n, a, sig, d_msgs, d_off, r, sl, kl);
    ctx->launches++;
    CU(cudaGetLastError());
    return dev_ed25519_verify(ctx, d, (const u32*)a, r, sl, kl, n, ok, s);
}

int dev_ed25519_verify(ecb_ctx* ctx, DevCtx& d, const u32* a, const u32* r, const u32* sl, const u32* kl, size_t n,
                       unsigned char* ok, cudaStream_t s) {
    TRY(dev_ed25519_table(ctx, d));
    TRY(ensure(ctx, d.cur->planes, n * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 8 * sizeof(u32)));
    unsigned g = persistent_grid(d, k_ed25519_verify, n);
    TRY(ensure(ctx, d.cur->scratch, (size_t)g * ECB_TPB * 8 * 32 * sizeof(u32)));
    TRY(reset_status(ctx, d, s));
    u32* planes = (u32*)d.cur->planes.p;
    prof_mark(ctx, d, s, 0);
    k_ed25519_verify<<<g, ECB_TPB, 0, s>>>(n, a, sl, kl, d.ed_table, d.ed_w, d.ed_nwin, d.ed_stride, (u32*)d.cur->scratch.p, planes, ok, d.cur->d_status);
    ctx->launches++;
    CU(cudaGetLastError());
    prof_mark(ctx, d, s, 1);
    // affine + encode_point + comparison with the signature's R bytes
    FinEdVerify fin{planes, n, r, ok};
    int rc = launch_batch_inv<F25519, FinEdVerify>(ctx, d, n, planes, (u32*)d.cur->pf.p, fin, s);
    prof_mark(ctx, d, s, 2);
    return rc;
}

// ---- key generation and signing (ed25519.rs:61-110).  Not constant-time. ------------
static __global__ void __launch_bounds__(ECB_TPB) k_ed25519_expand(size_t n, const unsigned char* seeds, u32* a_out, u32* prefix_out) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) ed25519_expand_body(idx, seeds, a_out, prefix_out);
}
static __global__ void __launch_bounds__(ECB_TPB) k_ed25519_sign_nonce(size_t n, const u32* prefix, const unsigned char* msgs,
                                                                 const unsigned long long* off, u32* r_out) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) ed25519_sign_nonce_body(idx, prefix, msgs, off, r_out);
}
static __global__ void __launch_bounds__(ECB_TPB) k_ed25519_sign_finish(size_t n, unsigned char* sig, const unsigned char* a_pub,
                                                                  const unsigned char* msgs, const unsigned long long* off,
                                                                  const u32* a_red, const u32* r_red) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) ed25519_sign_finish_body(idx, sig, a_pub, msgs, off, a_red, r_red);
}
// SecretKey::public_key (ed25519.rs:81 public_from_seed): seeds n x 32 -> encode_point(a B) n x 32
// The working buffers of key generation and signing hold secrets (expanded scalar a, prefix, nonce r, and the
// projective planes of a B / r B): they are cleared on the stream before the call's last kernel retires.
static int wipe_secrets(ecb_ctx* ctx, DevCtx& d, size_t n, cudaStream_t s) {
    if (d.cur->aux.p) CU(cudaMemsetAsync(d.cur->aux.p, 0, d.cur->aux.cap < n * 4 * 32 ? d.cur->aux.cap : n * 4 * 32, s));
    if (d.cur->planes.p) CU(cudaMemsetAsync(d.cur->planes.p, 0, d.cur->planes.cap < n * 96 ? d.cur->planes.cap : n * 96, s));
    if (d.cur->pf.p) CU(cudaMemsetAsync(d.cur->pf.p, 0, d.cur->pf.cap < n * 32 ? d.cur->pf.cap : n * 32, s));
    return ECB_OK;
}
static int ed_secret_mul_base(ecb_ctx* ctx, DevCtx& d, bool ct, const u32* k, size_t n, u32* out, cudaStream_t s, size_t stride_words) {
    return ct ? dev_ed25519_mul_base_ct(ctx, d, k, n, out, true, s, stride_words) : dev_ed25519_mul_base(ctx, d, k, n, out, true, s, stride_words);
}
int dev_ed25519_public_from_seed(ecb_ctx* ctx, DevCtx& d, const unsigned char* d_seeds, size_t n, u32* d_pub, cudaStream_t s, bool ct) {
    TRY(ensure(ctx, d.cur->aux, n * 4 * 32));
    u32* a = (u32*)d.cur->aux.p;
    u32* prefix = a + n * 8;
    k_ed25519_expand<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_seeds, a, prefix);
    ctx->launches++;
    CU(cudaGetLastError());
    TRY(ed_secret_mul_base(ctx, d, ct, a, n, d_pub, s, 0));
    return wipe_secrets(ctx, d, n, s);
}
// Keypair::sign / sign_with_public (ed25519.rs:94): d_pub may be null (SecretKey::sign, :112: A is derived first)
int dev_ed25519_sign(ecb_ctx* ctx, DevCtx& d, const unsigned char* d_seeds, const unsigned char* d_pub, const unsigned char* d_msgs,
                     const unsigned long long* d_off, size_t n, unsigned char* d_sig, cudaStream_t s, bool ct) {
    TRY(ensure(ctx, d.cur->aux, n * 4 * 32));
    u32* a = (u32*)d.cur->aux.p;
    u32* prefix = a + n * 8;
    u32* r = prefix + n * 8;
    u32* pub_tmp = r + n * 8;
    k_ed25519_expand<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_seeds, a, prefix);
    ctx->launches++;
    CU(cudaGetLastError());
    if (!d_pub) {
        TRY(ed_secret_mul_base(ctx, d, ct, a, n, pub_tmp, s, 0));
        d_pub = (const unsigned char*)pub_tmp;
    }
    k_ed25519_sign_nonce<<<grid_for(n), ECB_TPB, 0, s>>>(n, prefix, d_msgs, d_off, r);
    ctx->launches++;
    CU(cudaGetLastError());
    TRY(ed_secret_mul_base(ctx, d, ct, r, n, (u32*)d_sig, s, 16));   // R = encode_point(r B) into bytes [0, 32) of each signature
    k_ed25519_sign_finish<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_sig, d_pub, d_msgs, d_off, a, r);
    ctx->launches++;
    CU(cudaGetLastError());
    return wipe_secrets(ctx, d, n, s);
}

// ---- ristretto255 (ristretto.cuh; src/curve/curve25519/ristretto255.rs) ------------------------------------
static __global__ void __launch_bounds__(ECB_TPB) k_ristretto255_decompress(size_t n, const u32* enc, u32* out_xy, unsigned char* ok) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) ristretto255_decompress_body(idx, enc, out_xy, ok);
}
static __global__ void __launch_bounds__(ECB_TPB) k_ristretto255_compress(size_t n, const u32* xy, u32* enc) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) ristretto255_compress_body(idx, xy, enc);
}
int dev_ristretto255_decompress(ecb_ctx* ctx, DevCtx& d, const u32* d_enc, size_t n, u32* d_out, unsigned char* d_ok, cudaStream_t s) {
    (void)d;
    k_ristretto255_decompress<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_enc, d_out, d_ok);
    ctx->launches++;
    CU(cudaGetLastError());
    return ECB_OK;
}
int dev_ristretto255_compress(ecb_ctx* ctx, DevCtx& d, const u32* d_xy, size_t n, u32* d_enc, cudaStream_t s) {
    (void)d;
    k_ristretto255_compress<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_xy, d_enc);
    ctx->launches++;
    CU(cudaGetLastError());
    return ECB_OK;
}
// RistrettoPoint::scale (ristretto255.rs:152): decode -> k * P on edwards25519 -> encode.  An invalid encoding decodes
// to (0, 0), which the Edwards kernel reports as a bad point with its index.
int dev_ristretto255_mul(ecb_ctx* ctx, DevCtx& d, const u32* d_k, const u32* d_enc, size_t n, u32* d_out, cudaStream_t s) {
    TRY(ensure(ctx, d.cur->aux, n * 64));
    u32* xy = (u32*)d.cur->aux.p;
    TRY(dev_ristretto255_decompress(ctx, d, d_enc, n, xy, nullptr, s));
    TRY(dev_ed25519_mul(ctx, d, d_k, xy, n, xy, s));
    return dev_ristretto255_compress(ctx, d, xy, n, d_out, s);
}
// RistrettoPoint::mul_base (ristretto255.rs:157)
int dev_ristretto255_mul_base(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, cudaStream_t s) {
    TRY(ensure(ctx, d.cur->aux, n * 64));
    u32* xy = (u32*)d.cur->aux.p;
    TRY(dev_ed25519_mul_base(ctx, d, d_k, n, xy, false, s));
    return dev_ristretto255_compress(ctx, d, xy, n, d_out, s);
}
