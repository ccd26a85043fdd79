// host_ctx.h — per-device context, error plumbing and launch helpers shared by the translation
// units of libeccbatch (one TU per kernel family so they compile in parallel).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/eccbatch.h"
#include "limb.cuh"

#define ECB_TPB 128
using ecb::u32;

// =======================================================================================
// context
// =======================================================================================
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// One pipeline slot: a stream with its own staging and work buffers.  The host entry points cut a
// device's slice into chunks and rotate them over ECB_NSLOT slots, so the H2D copy of chunk i+1, the
// kernels of chunk i and the D2H copy of chunk i-1 overlap (two copy engines + SMs); measured e2e at n = 2^20:
// 3 slots 491 M/s, 4 slots 500 M/s (Ed25519 mul_base).
#define ECB_NSLOT 4
struct Slot {
    cudaStream_t stream = nullptr;
    DevBuf planes, pf, scratch, aux, in[4], out[3];
    unsigned long long* d_status = nullptr;
    unsigned long long* h_status = nullptr;  // pinned
    bool busy = false;
    cudaEvent_t ev_join = nullptr;            // device-resident fork/join (see dev_forkjoin)
    cudaStream_t hi = nullptr;                // high-priority side stream for the batch-inversion kernel
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    size_t c0 = 0;                            // first element of the chunk in flight
};

struct DevCtx {
    int dev = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;            // = slots[0].stream: table builds, probes
    Slot slots[ECB_NSLOT];
    cudaEvent_t ev_fork = nullptr;
    size_t inv_fill_per_sm = 256;             // lower bound on batch-inversion threads per SM (option "inv_fill_per_sm")
    int dev_slots_used = 1;                   // slots whose status words the last *_dev call wrote
    int inv_block = 1;                        // batch inversion, one safegcd per block instead of per thread: 0 never, 1 where measured faster (2^255-19), 2 every field (option "inv_block")
    size_t inv_per_thread = 32;               // batch-inversion chain length (option "inv_per_thread"; 8: 0.31 ms, 16: 0.24, 32: 0.20, 64: 0.20 at n = 2^20)
    Slot* cur = &slots[0];                    // slot whose buffers the dev_* functions use (calls are serialised by `mu`)
    u32* ed_table = nullptr;
    int ed_w = 0, ed_nwin = 0, ed_stride = 24;  // comb width, windows, words between entries (24 packed, 32 = 128-byte aligned)
    u32* ed_ct_table = nullptr;               // constant-time comb (ct.cuh): W = 4, 64 windows x 8 entries, 48 KB
    u32* wei_ct_table[4] = {nullptr, nullptr, nullptr, nullptr};   // constant-time generator combs of the Weierstrass curves (W = 4)
    u32* wei_table[4] = {nullptr, nullptr, nullptr, nullptr};  // generator combs of p256r1, p384r1, bls12_381 G1, p256k1
    int wei_w[4] = {0, 0, 0, 0}, wei_nwin[4] = {0, 0, 0, 0};
    DevBuf trace;                             // option "trace": per-block timestamps of the last fused launch
    size_t trace_blocks = 0;
    std::mutex mu;
    // optional per-kernel timing (ecb_set_option "profile"): events recorded on the launching stream
    // around the scalar-mult kernel(s) [a,b] and the batch-inversion finisher [b,c] of every call
    struct ProfRec { cudaEvent_t a, b, c; };
    std::vector<ProfRec> prof;
};

struct ecb_ctx {
    std::vector<DevCtx*> devs;
    std::string err;
    std::mutex err_mu;
    long opt_wei_w[4] = {0, 0, 0, 0};  // generator comb widths (p256r1, p384r1, bls12_381 g1, p256k1); 0 = by free memory: 24 / 22 / 24 (11 / 18 / 11 windows, 5.9 / 3.6 / 8.9 GB) above 100 GB free, else 20 / 18 / 20
    // Ed25519 comb width; 0 = pick by free device memory (24: 11 windows, 8.9 GB table; 20: 13 windows, 0.65 GB;
    // 16: 16 windows, 50 MB).  Measured at n = 2^20: w=16 772 M/s, 20 907, 22 969, 24 1041 M/s — the kernel is
    // integer-pipe-bound, so time follows the window count; the random table reads (96 B per window) stay
    // far below HBM bandwidth.
    long opt_ed_w = 0;
    // words between comb entries: 32 (default: one 96-byte entry per 128-byte line) or 24 (packed).  Measured at W = 24,
    // n = 2^20 (profiles/r02_l_*): same kernel time (0.79 vs 0.80 ms), DRAM reads 1.37 GB vs 2.03 GB per launch against a
    // minimum of 1.11 GB — packed entries straddle DRAM atoms; the aligned table is 11.8 GB instead of 8.9 GB.
    long opt_ed_stride = 32;
    long opt_ed_fused = 1;                    // small-batch fused kernel (fused.cuh): 0 never, 1 when the batch fits one wave, 2 always
    long opt_ed_lanes = 0;                    // lanes per scalar in the fused kernel: 0 = by batch size, or 1 / 2 / 4 / 8
    size_t opt_chunk = 189440;  // elements per pipeline chunk = 148 SMs x 1280 (ECB_NSLOT chunks in flight per device; measured best of 2^16..2^19)
    long opt_profile = 0;
    long opt_bls_glv = 0;                     // 1: ecb_wei_mul on bls12_381_g1 takes its inputs to be in G1 and uses the endomorphism (kernels3.cuh)
    long opt_trace = 0;                       // 1: the fused kernels record per-block phase timestamps (measurement only)
    long opt_inv_hi = 1;                      // run batch inversions on the slot's high-priority side stream
    long opt_ramp = 2;                        // pipeline chunk schedule: this many halvings of the chunk size at both ends of a batch (option "ramp")
    long opt_dev_split = 0;                   // 1: split large device-resident batches over the slot streams (measured: no gain, the
                                              // shorter inversion chains cost what the overlap saves; kept as an option)
    std::atomic<unsigned long long> launches{0};
    // set while a *_dev entry point runs: those enqueue only — no allocation, no table build, no synchronisation.
    // Anything missing makes the call fail with ECB_ERR_NOT_READY; ecb_warm() provides it beforehand.
    bool no_alloc = false;
};

static inline int set_err(ecb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) {
        std::lock_guard<std::mutex> g(ctx->err_mu);
        ctx->err = msg;
    }
    return code;
}
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? ECB_ERR_OOM : ECB_ERR_CUDA;              \
            return set_err(ctx, code_, std::string(#call) + ": " + cudaGetErrorString(e_));          \
        }                                                                                            \
    } while (0)

static inline int ensure(ecb_ctx* ctx, DevBuf& b, size_t bytes) {
    if (b.cap >= bytes) return ECB_OK;
    if (ctx->no_alloc) return set_err(ctx, ECB_ERR_NOT_READY, "a *_dev call needs a larger work buffer: call ecb_warm(ctx, op, curve, max_n) first");
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8;
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    return ECB_OK;
}
#define TRY(x)                 \
    do {                       \
        int r_ = (x);          \
        if (r_ != ECB_OK) return r_; \
    } while (0)

static inline unsigned grid_for(size_t n) { return (unsigned)((n + ECB_TPB - 1) / ECB_TPB); }

// number of threads for the batch inversion: ~16 elements per thread, but never fewer threads
// than fill the machine once
static inline size_t inv_threads(const DevCtx& d, size_t n) {
    size_t per = d.inv_per_thread ? d.inv_per_thread : 32;
    size_t T = (n + per - 1) / per;
    size_t fill = (size_t)d.sm_count * (d.inv_fill_per_sm ? d.inv_fill_per_sm : 256);
    if (T < fill) T = fill;
    if (T > n) T = n;
    return T ? T : 1;
}

template <class K>
static inline unsigned persistent_grid(const DevCtx& d, K kernel, size_t n) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ECB_TPB, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    size_t g = (size_t)d.sm_count * per_sm;
    size_t need = (n + ECB_TPB - 1) / ECB_TPB;
    if (g > need) g = need;
    return (unsigned)(g ? g : 1);
}

// d_status[0]: the per-batch status word (~0 = no error); d_status[1]: the work counter of the persistent kernels
// (tu_common.cuh warp_next_chunk), which starts at ~0 = "-1" so that ONE memset arms both
static inline int reset_status(ecb_ctx* ctx, DevCtx& d, cudaStream_t s) {
    CU(cudaMemsetAsync(d.cur->d_status, 0xff, 2 * sizeof(unsigned long long), s));
    return ECB_OK;
}


// profiling marks: call prof_begin before the main kernel(s), prof_mid between them and the
// finisher, prof_end after it.  No-ops unless the "profile" option is set.
static inline void prof_mark(ecb_ctx* ctx, DevCtx& d, cudaStream_t s, int which) {
    if (!ctx->opt_profile) return;
    if (which == 0) {
        DevCtx::ProfRec r;
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventCreate(&r.c);
        d.prof.push_back(r);
        cudaEventRecord(r.a, s);
    } else if (!d.prof.empty()) {
        cudaEventRecord(which == 1 ? d.prof.back().b : d.prof.back().c, s);
    }
}
