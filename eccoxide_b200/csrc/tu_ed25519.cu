// tu_ed25519.cu — edwards25519 kernels: fixed-base comb (+ table builder), variable base, verify.
#include "tu_common.cuh"
#include "dev_ops.h"

static __global__ void __launch_bounds__(ECB_TPB, 5) k_ed25519_mul_base(size_t n, const u32* scalars, const u32* table, int W,
                                                               int nwin, u32* planes, unsigned long long* status) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) ed25519_mul_base_body<false>(idx, n, scalars, table, W, nwin, planes, status);
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
    size_t T = (size_t)gridDim.x * ECB_TPB, t = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    u32* tbl = scratch + t * (8 * 32);
    for (size_t idx = t; idx < n; idx += T) ed25519_mul_body(idx, n, scalars, points, tbl, planes, status);
}
static __global__ void __launch_bounds__(ECB_TPB, 3) k_ed25519_verify(size_t n, const u32* a_enc, const u32* s_le, const u32* k_le,
                                                             const u32* table, int W, int nwin, u32* scratch, u32* planes,
                                                             unsigned char* ok) {
    size_t T = (size_t)gridDim.x * ECB_TPB, t = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    u32* tbl = scratch + t * (8 * 32);
    for (size_t idx = t; idx < n; idx += T) ed25519_verify_body(idx, n, a_enc, s_le, k_le, table, W, nwin, tbl, planes, ok);
}

// comb width actually used on this device: the option, or (option 0) the widest whose table and
// transient build buffers (~2.5x the table) fit comfortably in the free memory
static int ed_pick_w(ecb_ctx* ctx) {
    if (ctx->opt_ed_w) return (int)ctx->opt_ed_w;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 16;
    if (free_b > ((size_t)110 << 30)) return 26;   // 10 windows, 32 GB table (+ 43 GB while it is built)
    if (free_b > ((size_t)40 << 30)) return 24;
    if (free_b > ((size_t)8 << 30)) return 20;
    return 16;
}
static int ed25519_build_table_w(ecb_ctx* ctx, DevCtx& d, int W);
// In automatic mode (option 0) a width whose table or build buffers cannot be allocated falls back
// to the next narrower even width instead of failing the call.
int dev_ed25519_build_table(ecb_ctx* ctx, DevCtx& d, int W) {
    for (;;) {
        int rc = ed25519_build_table_w(ctx, d, W);
        if (rc == ECB_OK || ctx->opt_ed_w != 0 || W <= 16) return rc;
        (void)cudaGetLastError();
        if (d.ed_table) cudaFree(d.ed_table);
        d.ed_table = nullptr;
        for (DevBuf* b : {&d.cur->planes, &d.cur->pf}) {
            if (b->p) cudaFree(b->p);
            b->p = nullptr;
            b->cap = 0;
        }
        W -= 2;
    }
}
static int ed25519_build_table_w(ecb_ctx* ctx, DevCtx& d, int W) {
    int nwin = (254 + W - 1) / W;
    size_t ntab = (size_t)nwin << (W - 1);
    if (d.ed_table) CU(cudaFree(d.ed_table));
    d.ed_table = nullptr;
    CU(cudaMalloc(&d.ed_table, ntab * 24 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->planes, ntab * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, ntab * 8 * sizeof(u32)));
    u32* bases = nullptr;   // 2^(W i) * B per window, cached form (32 words each)
    CU(cudaMalloc(&bases, (size_t)nwin * 32 * sizeof(u32)));
    k_ed25519_window_bases<<<grid_for((size_t)nwin), ECB_TPB, 0, d.stream>>>(nwin, W, bases);
    k_ed25519_table_points<<<grid_for(ntab), ECB_TPB, 0, d.stream>>>(ntab, W, nwin, bases, (u32*)d.cur->planes.p);
    ctx->launches += 2;
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        cudaFree(bases);
        CU(le);
    }
    FinEdNiels fin{(const u32*)d.cur->planes.p, ntab, d.ed_table};
    TRY((launch_batch_inv<F25519, FinEdNiels>(ctx, d, ntab, (const u32*)d.cur->planes.p, (u32*)d.cur->pf.p, fin, d.stream)));
    cudaError_t se = cudaStreamSynchronize(d.stream);
    cudaFree(bases);
    CU(se);
    if (ntab * 96 > ((size_t)256 << 20)) {  // give the build buffers of a large table back
        for (DevBuf* b : {&d.cur->planes, &d.cur->pf}) {
            if (b->p) CU(cudaFree(b->p));
            b->p = nullptr;
            b->cap = 0;
        }
    }
    d.ed_w = W;
    d.ed_nwin = nwin;
    return ECB_OK;
}

int dev_ed25519_mul_base(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, bool compressed, cudaStream_t s,
                         size_t enc_stride_words) {
    if (!d.ed_table || (ctx->opt_ed_w && d.ed_w != (int)ctx->opt_ed_w)) TRY(dev_ed25519_build_table(ctx, d, ed_pick_w(ctx)));
    TRY(ensure(ctx, d.cur->planes, n * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 8 * sizeof(u32)));
    TRY(reset_status(ctx, d, s));
    u32* planes = (u32*)d.cur->planes.p;
    prof_mark(ctx, d, s, 0);
    k_ed25519_mul_base<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_k, d.ed_table, d.ed_w, d.ed_nwin, planes, d.cur->d_status);
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

static __global__ void __launch_bounds__(ECB_TPB, 5) k_x25519_base(size_t n, const u32* scalars, const u32* table, int W, int nwin,
                                                          u32* planes) {
    size_t idx = (size_t)blockIdx.x * ECB_TPB + threadIdx.x;
    if (idx < n) x25519_base_body(idx, n, scalars, table, W, nwin, planes);
}
int dev_x25519_base(ecb_ctx* ctx, DevCtx& d, const u32* d_k, size_t n, u32* d_out, cudaStream_t s) {
    if (!d.ed_table || (ctx->opt_ed_w && d.ed_w != (int)ctx->opt_ed_w)) TRY(dev_ed25519_build_table(ctx, d, ed_pick_w(ctx)));
    if (d.ed_w * d.ed_nwin < 256) return set_err(ctx, ECB_ERR_INVALID_ARG, "x25519_base needs a comb covering 256 bits (ed25519_comb_w)");
    TRY(ensure(ctx, d.cur->planes, n * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 8 * sizeof(u32)));
    TRY(reset_status(ctx, d, s));
    u32* planes = (u32*)d.cur->planes.p;
    prof_mark(ctx, d, s, 0);
    k_x25519_base<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_k, d.ed_table, d.ed_w, d.ed_nwin, planes);
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
    if (!d.ed_table || (ctx->opt_ed_w && d.ed_w != (int)ctx->opt_ed_w)) TRY(dev_ed25519_build_table(ctx, d, ed_pick_w(ctx)));
    TRY(ensure(ctx, d.cur->planes, n * 3 * 8 * sizeof(u32)));
    TRY(ensure(ctx, d.cur->pf, n * 8 * sizeof(u32)));
    unsigned g = persistent_grid(d, k_ed25519_verify, n);
    TRY(ensure(ctx, d.cur->scratch, (size_t)g * ECB_TPB * 8 * 32 * sizeof(u32)));
    TRY(reset_status(ctx, d, s));
    u32* planes = (u32*)d.cur->planes.p;
    prof_mark(ctx, d, s, 0);
    k_ed25519_verify<<<g, ECB_TPB, 0, s>>>(n, a, sl, kl, d.ed_table, d.ed_w, d.ed_nwin, (u32*)d.cur->scratch.p, planes, ok);
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
int dev_ed25519_public_from_seed(ecb_ctx* ctx, DevCtx& d, const unsigned char* d_seeds, size_t n, u32* d_pub, cudaStream_t s) {
    TRY(ensure(ctx, d.cur->aux, n * 4 * 32));
    u32* a = (u32*)d.cur->aux.p;
    u32* prefix = a + n * 8;
    k_ed25519_expand<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_seeds, a, prefix);
    ctx->launches++;
    CU(cudaGetLastError());
    return dev_ed25519_mul_base(ctx, d, a, n, d_pub, true, s, 0);
}
// Keypair::sign / sign_with_public (ed25519.rs:94): d_pub may be null (SecretKey::sign, :112: A is derived first)
int dev_ed25519_sign(ecb_ctx* ctx, DevCtx& d, const unsigned char* d_seeds, const unsigned char* d_pub, const unsigned char* d_msgs,
                     const unsigned long long* d_off, size_t n, unsigned char* d_sig, cudaStream_t s) {
    TRY(ensure(ctx, d.cur->aux, n * 4 * 32));
    u32* a = (u32*)d.cur->aux.p;
    u32* prefix = a + n * 8;
    u32* r = prefix + n * 8;
    u32* pub_tmp = r + n * 8;
    k_ed25519_expand<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_seeds, a, prefix);
    ctx->launches++;
    CU(cudaGetLastError());
    if (!d_pub) {
        TRY(dev_ed25519_mul_base(ctx, d, a, n, pub_tmp, true, s, 0));
        d_pub = (const unsigned char*)pub_tmp;
    }
    k_ed25519_sign_nonce<<<grid_for(n), ECB_TPB, 0, s>>>(n, prefix, d_msgs, d_off, r);
    ctx->launches++;
    CU(cudaGetLastError());
    TRY(dev_ed25519_mul_base(ctx, d, r, n, (u32*)d_sig, true, s, 16));   // R = encode_point(r B) into bytes [0, 32) of each signature
    k_ed25519_sign_finish<<<grid_for(n), ECB_TPB, 0, s>>>(n, d_sig, d_pub, d_msgs, d_off, a, r);
    ctx->launches++;
    CU(cudaGetLastError());
    return ECB_OK;
}
