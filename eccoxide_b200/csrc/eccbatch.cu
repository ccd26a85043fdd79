// eccbatch.cu — the C ABI of include/eccbatch.h: context life cycle, sharding of a batch over the
// devices of the context, chunking, host<->device copies.  All arithmetic lives in the kernel TUs
// (tu_*.cu); there is no host arithmetic path in this library.
#include "dev_ops.h"
using namespace ecb;

static int dev_wei_mul(ecb_ctx* ctx, DevCtx& d, int curve, const u32* d_k, const u32* d_p, const unsigned char* d_inf_in,
                       size_t n, u32* d_out, unsigned char* d_inf, cudaStream_t s) {
    switch (curve) {
        case ECB_CURVE_P256R1: return dev_wei_mul_p256(ctx, d, d_k, d_p, d_inf_in, n, d_out, d_inf, s);
        case ECB_CURVE_P384R1: return dev_wei_mul_p384(ctx, d, d_k, d_p, d_inf_in, n, d_out, d_inf, s);
        case ECB_CURVE_BLS12_381_G1: return dev_wei_mul_bls(ctx, d, d_k, d_p, d_inf_in, n, d_out, d_inf, s);
        case ECB_CURVE_P256K1: return dev_wei_mul_k256(ctx, d, d_k, d_p, d_inf_in, n, d_out, d_inf, s);
    }
    return set_err(ctx, ECB_ERR_INVALID_ARG, "unknown curve id");
}

// =======================================================================================
// C ABI
// =======================================================================================
extern "C" {

int ecb_init(const int* device_ids, int n_dev, ecb_ctx** out) {
    if (!out) return ECB_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return ECB_ERR_CUDA;  // no CPU fallback
    ecb_ctx* ctx = new ecb_ctx();
    int one = 0;
    if (!device_ids || n_dev <= 0) {
        device_ids = &one;
        n_dev = 1;
    }
    for (int i = 0; i < n_dev; i++) {
        if (device_ids[i] < 0 || device_ids[i] >= count) {
            ecb_destroy(ctx);
            return ECB_ERR_INVALID_ARG;
        }
        DevCtx* d = new DevCtx();
        d->dev = device_ids[i];
        ctx->devs.push_back(d);
        bool ok = cudaSetDevice(d->dev) == cudaSuccess &&
                  cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, d->dev) == cudaSuccess;
        for (int s = 0; ok && s < ECB_NSLOT; s++) {
            Slot& sl = d->slots[s];
            int lo_p = 0, hi_p = 0;
            cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
            ok = cudaStreamCreateWithPriority(&sl.stream, cudaStreamNonBlocking, lo_p) == cudaSuccess &&
                 cudaStreamCreateWithPriority(&sl.hi, cudaStreamNonBlocking, hi_p) == cudaSuccess &&
                 cudaEventCreateWithFlags(&sl.ev_a, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&sl.ev_b, cudaEventDisableTiming) == cudaSuccess &&
                 cudaMalloc(&sl.d_status, 2 * sizeof(unsigned long long)) == cudaSuccess &&
                 cudaMallocHost(&sl.h_status, sizeof(unsigned long long)) == cudaSuccess &&
                 cudaMemset(sl.d_status, 0xff, 2 * sizeof(unsigned long long)) == cudaSuccess;
        }
        if (!ok) {
            ecb_destroy(ctx);
            return ECB_ERR_CUDA;
        }
        d->stream = d->slots[0].stream;
    }
    *out = ctx;
    return ECB_OK;
}

void ecb_destroy(ecb_ctx* ctx) {
    if (!ctx) return;
    for (DevCtx* d : ctx->devs) {
        cudaSetDevice(d->dev);
        for (Slot& sl : d->slots) {
            if (sl.stream) cudaStreamSynchronize(sl.stream);
            DevBuf* bufs[] = {&sl.planes, &sl.pf, &sl.scratch, &sl.aux, &sl.in[0], &sl.in[1], &sl.in[2], &sl.in[3], &sl.out[0], &sl.out[1], &sl.out[2]};
            for (DevBuf* b : bufs)
                if (b->p) {
                    cudaMemset(b->p, 0, b->cap);   // work buffers may hold key material of signing / key-generation calls
                    cudaFree(b->p);
                }
            if (sl.d_status) cudaFree(sl.d_status);
            if (sl.h_status) cudaFreeHost(sl.h_status);
            if (sl.ev_join) cudaEventDestroy(sl.ev_join);
            if (sl.ev_a) cudaEventDestroy(sl.ev_a);
            if (sl.ev_b) cudaEventDestroy(sl.ev_b);
            if (sl.hi) cudaStreamDestroy(sl.hi);
            if (sl.stream) cudaStreamDestroy(sl.stream);
        }
        for (auto& r : d->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); cudaEventDestroy(r.c); }
        if (d->ed_table) cudaFree(d->ed_table);
        if (d->trace.p) cudaFree(d->trace.p);
        if (d->ed_ct_table) cudaFree(d->ed_ct_table);
        for (u32* t : d->wei_ct_table) if (t) cudaFree(t);
        if (d->ev_fork) cudaEventDestroy(d->ev_fork);
        for (u32* t : d->wei_table)
            if (t) cudaFree(t);
        delete d;
    }
    delete ctx;
}

const char* ecb_last_error(ecb_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (no CUDA device?)"; }
int ecb_device_count(ecb_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }
unsigned long long ecb_launch_count(ecb_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int ecb_set_option(ecb_ctx* ctx, const char* key, long value) {
    if (!ctx || !key) return ECB_ERR_INVALID_ARG;
    if (!strcmp(key, "ed25519_comb_w")) {
        if (value != 0 && (value < 4 || value > 26)) return set_err(ctx, ECB_ERR_INVALID_ARG, "ed25519_comb_w must be 0 or in 4..26");
        ctx->opt_ed_w = value;
        return ECB_OK;
    }
    if (!strcmp(key, "ed25519_entry_stride")) {
        if (value != 24 && value != 32) return set_err(ctx, ECB_ERR_INVALID_ARG, "ed25519_entry_stride must be 24 or 32 (32-bit words)");
        ctx->opt_ed_stride = value;
        return ECB_OK;
    }
    if (!strcmp(key, "ed25519_fused")) {
        if (value < 0 || value > 2) return set_err(ctx, ECB_ERR_INVALID_ARG, "ed25519_fused must be 0, 1 or 2");
        ctx->opt_ed_fused = value;
        return ECB_OK;
    }
    if (!strcmp(key, "ed25519_lanes")) {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return set_err(ctx, ECB_ERR_INVALID_ARG, "ed25519_lanes must be 0, 1, 2, 4 or 8");
        ctx->opt_ed_lanes = value;
        return ECB_OK;
    }
    if (!strcmp(key, "p256r1_comb_w") || !strcmp(key, "p384r1_comb_w") || !strcmp(key, "bls12_381_g1_comb_w") || !strcmp(key, "p256k1_comb_w")) {
        if (value != 0 && (value < 4 || value > 24)) return set_err(ctx, ECB_ERR_INVALID_ARG, "comb width must be 0 or in 4..24");
        ctx->opt_wei_w[!strcmp(key, "p256k1_comb_w") ? 3 : (key[1] == '2' ? 0 : (key[1] == '3' ? 1 : 2))] = value;
        return ECB_OK;
    }
    if (!strcmp(key, "inv_per_thread")) {
        if (value < 1 || value > 4096) return set_err(ctx, ECB_ERR_INVALID_ARG, "inv_per_thread must be in 1..4096");
        for (DevCtx* d : ctx->devs) d->inv_per_thread = (size_t)value;
        return ECB_OK;
    }
    if (!strcmp(key, "inv_block")) {
        if (value < 0 || value > 2) return set_err(ctx, ECB_ERR_INVALID_ARG, "inv_block must be 0, 1 or 2");
        for (DevCtx* d : ctx->devs) d->inv_block = (int)value;
        return ECB_OK;
    }
    if (!strcmp(key, "inv_fill_per_sm")) {
        if (value < 32 || value > 2048) return set_err(ctx, ECB_ERR_INVALID_ARG, "inv_fill_per_sm must be in 32..2048");
        for (DevCtx* d : ctx->devs) d->inv_fill_per_sm = (size_t)value;
        return ECB_OK;
    }
    if (!strcmp(key, "chunk")) {
        if (value < 1) return set_err(ctx, ECB_ERR_INVALID_ARG, "chunk must be >= 1");
        ctx->opt_chunk = (size_t)value;
        return ECB_OK;
    }
    if (!strcmp(key, "ramp")) {
        if (value < 0 || value > 4) return set_err(ctx, ECB_ERR_INVALID_ARG, "ramp must be in 0..4");
        ctx->opt_ramp = value;
        return ECB_OK;
    }
    if (!strcmp(key, "inv_hi")) {
        ctx->opt_inv_hi = value ? 1 : 0;
        return ECB_OK;
    }
    if (!strcmp(key, "dev_split")) {
        ctx->opt_dev_split = value ? 1 : 0;
        return ECB_OK;
    }
    if (!strcmp(key, "trace")) {
        ctx->opt_trace = value ? 1 : 0;
        return ECB_OK;
    }
    if (!strcmp(key, "bls12_381_g1_glv")) {
        ctx->opt_bls_glv = value ? 1 : 0;
        return ECB_OK;
    }
    if (!strcmp(key, "profile")) {
        ctx->opt_profile = value ? 1 : 0;
        return ECB_OK;
    }
    return set_err(ctx, ECB_ERR_INVALID_ARG, std::string("unknown option ") + key);
}

void* ecb_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void ecb_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"

// ---- host-buffer plumbing: shard over devices, chunk, copy, run, copy back ---------------------
struct HostArg {
    const uint8_t* src;  // input (nullptr = absent)
    size_t elem;         // bytes per element
};
struct HostOut {
    uint8_t* dst;
    size_t elem;
};

// op(d, stream, din[], dout[], n) enqueues the kernels for one chunk
// On an error path a host entry point must not return while copies it enqueued still read or write
// the caller's buffers: drain every slot stream of the device and reset the slot state.
struct SlotDrain {
    DevCtx& d;
    bool armed = true;
    ~SlotDrain() {
        if (!armed) return;
        for (Slot& sl : d.slots) {
            if (sl.stream) cudaStreamSynchronize(sl.stream);
            if (sl.hi) cudaStreamSynchronize(sl.hi);
            sl.busy = false;
        }
        d.cur = &d.slots[0];
    }
};

// Chunk schedule of one device's slice [lo, hi): the first chunks ramp up from chunk / 2^ramp and the
// last ones ramp back down, so that the pipeline (copy in | kernels | copy out, ECB_NSLOT chunks in
// flight) fills and drains on small chunks: the first copy-in and the last copy-out, which nothing
// can overlap, shrink with the chunk they belong to.  The middle is cut into equal chunks <= chunk.
static void chunk_plan(size_t lo, size_t hi, size_t chunk, long ramp, std::vector<size_t>& bounds) {
    bounds.clear();
    bounds.push_back(lo);
    size_t len = hi - lo;
    std::vector<size_t> up;
    size_t used = 0;
    for (long r = ramp; r >= 1; r--) {
        size_t c = chunk >> r;
        if (c < 4096) continue;
        up.push_back(c);
        used += 2 * c;
    }
    if (used + chunk > len) {  // short batch: no ramp
        up.clear();
        used = 0;
    }
    size_t pos = lo;
    for (size_t c : up) bounds.push_back(pos += c);
    size_t mid = len - used;
    size_t nmid = (mid + chunk - 1) / chunk;
    if (up.empty() && ramp > 0) {  // short batch: still ECB_NSLOT chunks (>= 16384 elements each) to overlap copies and kernels
        size_t want = len / 16384 < (size_t)ECB_NSLOT ? len / 16384 : (size_t)ECB_NSLOT;
        if (nmid < want) nmid = want;
    }
    for (size_t i = 1; i <= nmid; i++) bounds.push_back(pos + mid * i / nmid);
    pos += mid;
    for (size_t i = up.size(); i-- > 0;) bounds.push_back(pos += up[i]);
}

template <class OP>
static int run_sharded(ecb_ctx* ctx, size_t n, const std::vector<HostArg>& ins, const std::vector<HostOut>& outs,
                       bool has_status, size_t* bad_index, OP op) {
    if (!ctx) return ECB_ERR_CUDA;
    if (bad_index) *bad_index = (size_t)-1;
    if (n == 0) return ECB_OK;
    int nd = (int)ctx->devs.size();
    std::vector<int> rc(nd, ECB_OK);
    std::vector<unsigned long long> bad(nd, ~0ull);
    auto worker = [&](int di) {
        DevCtx& d = *ctx->devs[di];
        size_t lo = n * (size_t)di / nd, hi = n * (size_t)(di + 1) / nd;
        if (lo == hi) return;
        std::lock_guard<std::mutex> g(d.mu);
        auto body = [&]() -> int {
            CU(cudaSetDevice(d.dev));
            SlotDrain drain{d};
            // retire the chunk in flight on `sl`: wait, then look at its status word
            auto retire = [&](Slot& sl) -> int {
                if (!sl.busy) return ECB_OK;
                sl.busy = false;
                CU(cudaStreamSynchronize(sl.stream));
                if (has_status && *sl.h_status != ~0ull) {
                    unsigned long long v = *sl.h_status;
                    unsigned long long w = (((v >> 8) + sl.c0) << 8) | (v & 0xff);
                    if (w < bad[di]) bad[di] = w;
                }
                return ECB_OK;
            };
            int rcb = ECB_OK;
            size_t ci = 0;
            std::vector<size_t> bounds;
            chunk_plan(lo, hi, ctx->opt_chunk, ctx->opt_ramp, bounds);
            for (; ci + 1 < bounds.size() && rcb == ECB_OK && bad[di] == ~0ull; ci++) {
                size_t c0 = bounds[ci], cn = bounds[ci + 1] - c0;
                Slot& sl = d.slots[ci % ECB_NSLOT];
                auto one = [&]() -> int {
                    TRY(retire(sl));
                    d.cur = &sl;
                    const void* din[4] = {nullptr, nullptr, nullptr, nullptr};
                    void* dout[3] = {nullptr, nullptr, nullptr};
                    for (size_t i = 0; i < ins.size(); i++) {
                        if (!ins[i].src) continue;
                        TRY(ensure(ctx, sl.in[i], cn * ins[i].elem));
                        CU(cudaMemcpyAsync(sl.in[i].p, ins[i].src + c0 * ins[i].elem, cn * ins[i].elem, cudaMemcpyHostToDevice, sl.stream));
                        din[i] = sl.in[i].p;
                    }
                    for (size_t i = 0; i < outs.size(); i++) {
                        if (!outs[i].dst) continue;
                        TRY(ensure(ctx, sl.out[i], cn * outs[i].elem));
                        dout[i] = sl.out[i].p;
                    }
                    TRY(op(d, sl.stream, din, dout, cn));
                    for (size_t i = 0; i < outs.size(); i++) {
                        if (!outs[i].dst) continue;
                        CU(cudaMemcpyAsync(outs[i].dst + c0 * outs[i].elem, sl.out[i].p, cn * outs[i].elem, cudaMemcpyDeviceToHost, sl.stream));
                    }
                    if (has_status) CU(cudaMemcpyAsync(sl.h_status, sl.d_status, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sl.stream));
                    sl.busy = true;
                    sl.c0 = c0;
                    return ECB_OK;
                };
                rcb = one();
            }
            for (Slot& sl : d.slots) {
                int r2 = retire(sl);
                if (rcb == ECB_OK) rcb = r2;
            }
            d.cur = &d.slots[0];
            drain.armed = rcb != ECB_OK;
            return rcb;
        };
        rc[di] = body();
    };
    if (nd == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nd; i++) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < nd; i++)
        if (rc[i] != ECB_OK) return rc[i];
    unsigned long long first = ~0ull;
    for (int i = 0; i < nd; i++)
        if (bad[i] < first) first = bad[i];
    if (first != ~0ull) {
        if (bad_index) *bad_index = (size_t)(first >> 8);
        int code = (first & 0xff) == ecb::ST_NONCANONICAL_SCALAR ? ECB_ERR_NONCANONICAL_SCALAR : ECB_ERR_POINT_NOT_ON_CURVE;
        return set_err(ctx, code, code == ECB_ERR_NONCANONICAL_SCALAR ? "non-canonical scalar" : "point not on curve");
    }
    return ECB_OK;
}

static int dev_wei_mul_base(ecb_ctx* ctx, DevCtx& d, int curve, const u32* d_k, size_t n, u32* d_out, unsigned char* d_inf,
                            cudaStream_t s) {
    switch (curve) {
        case ECB_CURVE_P256R1: return dev_wei_mul_base_p256(ctx, d, d_k, n, d_out, d_inf, s);
        case ECB_CURVE_P384R1: return dev_wei_mul_base_p384(ctx, d, d_k, n, d_out, d_inf, s);
        case ECB_CURVE_BLS12_381_G1: return dev_wei_mul_base_bls(ctx, d, d_k, n, d_out, d_inf, s);
        case ECB_CURVE_P256K1: return dev_wei_mul_base_k256(ctx, d, d_k, n, d_out, d_inf, s);
    }
    return set_err(ctx, ECB_ERR_INVALID_ARG, "unknown curve id");
}

// Large device-resident batches of the fixed-base operations are cut into ECB_NSLOT sub-batches that
// run on the context's slot streams, forked from and joined back to the caller's stream with events:
// the batch-inversion kernel of one sub-batch (a long dependent chain per thread, latency-bound)
// then overlaps the scalar-multiplication kernel of the next (integer-pipe-bound).
// op(first, count, stream) enqueues one sub-batch using the buffers of d.cur.
template <class OP>
static int dev_forkjoin(ecb_ctx* ctx, DevCtx& d, size_t n, cudaStream_t s, OP op) {
    const size_t min_split = (size_t)3 << 16;
    if (!ctx->opt_dev_split || n < min_split) {
        d.cur = &d.slots[0];
        d.slots[0].c0 = 0;
        d.dev_slots_used = 1;
        return op((size_t)0, n, s);
    }
    if (!d.ev_fork) CU(cudaEventCreateWithFlags(&d.ev_fork, cudaEventDisableTiming));
    CU(cudaEventRecord(d.ev_fork, s));
    int rc = ECB_OK;
    for (int k = 0; k < ECB_NSLOT && rc == ECB_OK; k++) {
        Slot& sl = d.slots[k];
        if (!sl.ev_join) CU(cudaEventCreateWithFlags(&sl.ev_join, cudaEventDisableTiming));
        CU(cudaStreamWaitEvent(sl.stream, d.ev_fork, 0));
        size_t lo = n * (size_t)k / ECB_NSLOT, hi = n * (size_t)(k + 1) / ECB_NSLOT;
        d.cur = &sl;
        sl.c0 = lo;
        rc = op(lo, hi - lo, sl.stream);
        CU(cudaEventRecord(sl.ev_join, sl.stream));
        CU(cudaStreamWaitEvent(s, sl.ev_join, 0));
    }
    d.cur = &d.slots[0];
    d.dev_slots_used = ECB_NSLOT;
    return rc;
}

static int curve_sizes(int curve, size_t& fb, size_t& sb) {
    switch (curve) {
        case ECB_CURVE_P256R1: fb = 32; sb = 32; return ECB_OK;
        case ECB_CURVE_P384R1: fb = 48; sb = 48; return ECB_OK;
        case ECB_CURVE_BLS12_381_G1: fb = 48; sb = 32; return ECB_OK;
        case ECB_CURVE_P256K1: fb = 32; sb = 32; return ECB_OK;
    }
    return ECB_ERR_INVALID_ARG;
}

extern "C" {

int ecb_ed25519_mul_base(ecb_ctx* ctx, const uint8_t* k_le, size_t n, uint8_t* xy_le, size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!k_le || !xy_le)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k_le, 32}}, {{xy_le, 64}}, true, bad_index,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** out, size_t cn) {
                           return dev_ed25519_mul_base(ctx, d, (const u32*)in[0], cn, (u32*)out[0], false, s);
                       });
}
int ecb_ed25519_mul_base_compressed(ecb_ctx* ctx, const uint8_t* k_le, size_t n, uint8_t* enc, size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!k_le || !enc)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k_le, 32}}, {{enc, 32}}, true, bad_index,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** out, size_t cn) {
                           return dev_ed25519_mul_base(ctx, d, (const u32*)in[0], cn, (u32*)out[0], true, s);
                       });
}
int ecb_ed25519_mul(ecb_ctx* ctx, const uint8_t* k_le, const uint8_t* xy_in, size_t n, uint8_t* xy_out, size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!k_le || !xy_in || !xy_out)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k_le, 32}, {xy_in, 64}}, {{xy_out, 64}}, true, bad_index,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** out, size_t cn) {
                           return dev_ed25519_mul(ctx, d, (const u32*)in[0], (const u32*)in[1], cn, (u32*)out[0], s);
                       });
}
int ecb_ed25519_verify_prehashed(ecb_ctx* ctx, const uint8_t* a_enc, const uint8_t* r_enc, const uint8_t* s_le,
                                 const uint8_t* k_le, size_t n, uint8_t* ok) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!a_enc || !r_enc || !s_le || !k_le || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{a_enc, 32}, {r_enc, 32}, {s_le, 32}, {k_le, 32}}, {{ok, 1}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** out, size_t cn) {
                           return dev_ed25519_verify(ctx, d, (const u32*)in[0], (const u32*)in[1], (const u32*)in[2],
                                                     (const u32*)in[3], cn, (unsigned char*)out[0], s);
                       });
}
int ecb_x25519(ecb_ctx* ctx, const uint8_t* k, const uint8_t* u, size_t n, uint8_t* out) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!k || !u || !out)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k, 32}, {u, 32}}, {{out, 32}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_x25519(ctx, d, (const u32*)in[0], (const u32*)in[1], cn, (u32*)o[0], s);
                       });
}
// Ed25519 verification on raw messages (ragged input): same device sharding and slot rotation as
// run_sharded, with the message bytes of a chunk copied as one contiguous range.
int ecb_ed25519_verify(ecb_ctx* ctx, const uint8_t* a_enc, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* sig,
                       size_t n, uint8_t* ok) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!a_enc || !msg_off || !sig || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    if (n == 0) return ECB_OK;
    for (size_t i = 0; i < n; i++)
        if (msg_off[i + 1] < msg_off[i]) return set_err(ctx, ECB_ERR_INVALID_ARG, "message offsets must be non-decreasing");
    if (msg_off[n] > msg_off[0] && !msgs) return set_err(ctx, ECB_ERR_INVALID_ARG, "null message buffer");
    int nd = (int)ctx->devs.size();
    std::vector<int> rc(nd, ECB_OK);
    auto worker = [&](int di) {
        DevCtx& d = *ctx->devs[di];
        size_t lo = n * (size_t)di / nd, hi = n * (size_t)(di + 1) / nd;
        if (lo == hi) return;
        std::lock_guard<std::mutex> g(d.mu);
        auto body = [&]() -> int {
            CU(cudaSetDevice(d.dev));
            SlotDrain drain{d};
            size_t ci = 0;
            for (size_t c0 = lo; c0 < hi; c0 += ctx->opt_chunk, ci++) {
                size_t cn = hi - c0 < ctx->opt_chunk ? hi - c0 : ctx->opt_chunk;
                Slot& sl = d.slots[ci % ECB_NSLOT];
                if (sl.busy) {
                    sl.busy = false;
                    CU(cudaStreamSynchronize(sl.stream));
                }
                d.cur = &sl;
                size_t mbytes = (size_t)(msg_off[c0 + cn] - msg_off[c0]);
                TRY(ensure(ctx, sl.in[0], cn * 32));
                TRY(ensure(ctx, sl.in[1], cn * 64));
                TRY(ensure(ctx, sl.in[2], (cn + 1) * sizeof(uint64_t)));
                TRY(ensure(ctx, sl.in[3], mbytes + 16));
                TRY(ensure(ctx, sl.out[0], cn));
                CU(cudaMemcpyAsync(sl.in[0].p, a_enc + c0 * 32, cn * 32, cudaMemcpyHostToDevice, sl.stream));
                CU(cudaMemcpyAsync(sl.in[1].p, sig + c0 * 64, cn * 64, cudaMemcpyHostToDevice, sl.stream));
                CU(cudaMemcpyAsync(sl.in[2].p, msg_off + c0, (cn + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, sl.stream));
                if (mbytes) CU(cudaMemcpyAsync(sl.in[3].p, msgs + msg_off[c0], mbytes, cudaMemcpyHostToDevice, sl.stream));
                // the offsets stay absolute: rebase the message pointer instead
                const unsigned char* d_msgs = (const unsigned char*)sl.in[3].p - msg_off[c0];
                TRY(dev_ed25519_verify_msgs(ctx, d, (const unsigned char*)sl.in[0].p, (const unsigned char*)sl.in[1].p, d_msgs,
                                            (const unsigned long long*)sl.in[2].p, cn, (unsigned char*)sl.out[0].p, sl.stream));
                CU(cudaMemcpyAsync(ok + c0, sl.out[0].p, cn, cudaMemcpyDeviceToHost, sl.stream));
                sl.busy = true;
            }
            for (Slot& sl : d.slots) {
                if (!sl.busy) continue;
                sl.busy = false;
                CU(cudaStreamSynchronize(sl.stream));
            }
            d.cur = &d.slots[0];
            drain.armed = false;
            return ECB_OK;
        };
        rc[di] = body();
    };
    if (nd == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nd; i++) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < nd; i++)
        if (rc[i] != ECB_OK) return rc[i];
    return ECB_OK;
}
static int ed25519_public_from_seed_impl(ecb_ctx* ctx, const uint8_t* seeds, size_t n, uint8_t* pub, bool ct) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!seeds || !pub)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{seeds, 32}}, {{pub, 32}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** out, size_t cn) {
                           return dev_ed25519_public_from_seed(ctx, d, (const unsigned char*)in[0], cn, (u32*)out[0], s, ct);
                       });
}
int ecb_ed25519_public_from_seed(ecb_ctx* ctx, const uint8_t* seeds, size_t n, uint8_t* pub) {
    return ed25519_public_from_seed_impl(ctx, seeds, n, pub, true);
}
int ecb_ed25519_public_from_seed_vartime(ecb_ctx* ctx, const uint8_t* seeds, size_t n, uint8_t* pub) {
    return ed25519_public_from_seed_impl(ctx, seeds, n, pub, false);
}
// Point::mul_base for SECRET scalars: the constant-time kernel (ct.cuh), same bytes as ecb_ed25519_mul_base
int ecb_ed25519_mul_base_ct(ecb_ctx* ctx, const uint8_t* k_le, size_t n, uint8_t* xy_le, size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!k_le || !xy_le)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k_le, 32}}, {{xy_le, 64}}, true, bad_index,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** out, size_t cn) {
                           return dev_ed25519_mul_base_ct(ctx, d, (const u32*)in[0], cn, (u32*)out[0], false, s);
                       });
}
// Ed25519 signing of raw messages (ragged input): the chunk loop of ecb_ed25519_verify with seeds and
// (optional) public keys in, signatures out.
static int ed25519_sign_impl(ecb_ctx* ctx, const uint8_t* seeds, const uint8_t* pub, const uint8_t* msgs, const uint64_t* msg_off, size_t n,
                             uint8_t* sig, bool ct);
int ecb_ed25519_sign(ecb_ctx* ctx, const uint8_t* seeds, const uint8_t* pub, const uint8_t* msgs, const uint64_t* msg_off, size_t n,
                     uint8_t* sig) {
    return ed25519_sign_impl(ctx, seeds, pub, msgs, msg_off, n, sig, true);
}
int ecb_ed25519_sign_vartime(ecb_ctx* ctx, const uint8_t* seeds, const uint8_t* pub, const uint8_t* msgs, const uint64_t* msg_off, size_t n,
                             uint8_t* sig) {
    return ed25519_sign_impl(ctx, seeds, pub, msgs, msg_off, n, sig, false);
}
static int ed25519_sign_impl(ecb_ctx* ctx, const uint8_t* seeds, const uint8_t* pub, const uint8_t* msgs, const uint64_t* msg_off, size_t n,
                             uint8_t* sig, bool ct) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!seeds || !msg_off || !sig)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    if (n == 0) return ECB_OK;
    for (size_t i = 0; i < n; i++)
        if (msg_off[i + 1] < msg_off[i]) return set_err(ctx, ECB_ERR_INVALID_ARG, "message offsets must be non-decreasing");
    if (msg_off[n] > msg_off[0] && !msgs) return set_err(ctx, ECB_ERR_INVALID_ARG, "null message buffer");
    int nd = (int)ctx->devs.size();
    std::vector<int> rc(nd, ECB_OK);
    auto worker = [&](int di) {
        DevCtx& d = *ctx->devs[di];
        size_t lo = n * (size_t)di / nd, hi = n * (size_t)(di + 1) / nd;
        if (lo == hi) return;
        std::lock_guard<std::mutex> g(d.mu);
        auto body = [&]() -> int {
            CU(cudaSetDevice(d.dev));
            SlotDrain drain{d};
            size_t ci = 0;
            for (size_t c0 = lo; c0 < hi; c0 += ctx->opt_chunk, ci++) {
                size_t cn = hi - c0 < ctx->opt_chunk ? hi - c0 : ctx->opt_chunk;
                Slot& sl = d.slots[ci % ECB_NSLOT];
                if (sl.busy) {
                    sl.busy = false;
                    CU(cudaStreamSynchronize(sl.stream));
                }
                d.cur = &sl;
                size_t mbytes = (size_t)(msg_off[c0 + cn] - msg_off[c0]);
                TRY(ensure(ctx, sl.in[0], cn * 32));
                if (pub) TRY(ensure(ctx, sl.in[1], cn * 32));
                TRY(ensure(ctx, sl.in[2], (cn + 1) * sizeof(uint64_t)));
                TRY(ensure(ctx, sl.in[3], mbytes + 16));
                TRY(ensure(ctx, sl.out[0], cn * 64));
                CU(cudaMemcpyAsync(sl.in[0].p, seeds + c0 * 32, cn * 32, cudaMemcpyHostToDevice, sl.stream));
                if (pub) CU(cudaMemcpyAsync(sl.in[1].p, pub + c0 * 32, cn * 32, cudaMemcpyHostToDevice, sl.stream));
                CU(cudaMemcpyAsync(sl.in[2].p, msg_off + c0, (cn + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, sl.stream));
                if (mbytes) CU(cudaMemcpyAsync(sl.in[3].p, msgs + msg_off[c0], mbytes, cudaMemcpyHostToDevice, sl.stream));
                const unsigned char* d_msgs = (const unsigned char*)sl.in[3].p - msg_off[c0];
                TRY(dev_ed25519_sign(ctx, d, (const unsigned char*)sl.in[0].p, pub ? (const unsigned char*)sl.in[1].p : nullptr, d_msgs,
                                     (const unsigned long long*)sl.in[2].p, cn, (unsigned char*)sl.out[0].p, sl.stream, ct));
                CU(cudaMemcpyAsync(sig + c0 * 64, sl.out[0].p, cn * 64, cudaMemcpyDeviceToHost, sl.stream));
                sl.busy = true;
            }
            for (Slot& sl : d.slots) {
                if (!sl.busy) continue;
                sl.busy = false;
                CU(cudaStreamSynchronize(sl.stream));
            }
            d.cur = &d.slots[0];
            drain.armed = false;
            return ECB_OK;
        };
        rc[di] = body();
    };
    if (nd == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nd; i++) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < nd; i++)
        if (rc[i] != ECB_OK) return rc[i];
    return ECB_OK;
}
// ECDSA verification on raw messages (ragged input); layout of a chunk in the slot buffers:
// in[0] = Q, in[1] = r || s, in[2] = offsets, aux2 (in[3]) = z, message bytes in `scratch2` (out[1]).
int ecb_ecdsa_verify(ecb_ctx* ctx, int curve, int hash, const uint8_t* q_xy, const uint8_t* msgs, const uint64_t* msg_off,
                     const uint8_t* rs_be, size_t n, uint8_t* ok, size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    if (bad_index) *bad_index = (size_t)-1;
    size_t fb, sb;
    if (curve_sizes(curve, fb, sb) || (curve != ECB_CURVE_P256R1 && curve != ECB_CURVE_P384R1)) return set_err(ctx, ECB_ERR_INVALID_ARG, "ECDSA is defined for p256r1/p384r1 only");
    if (hash != 256 && hash != 384 && hash != 512) return set_err(ctx, ECB_ERR_INVALID_ARG, "hash must be 256, 384 or 512 (SHA-2)");
    if (n && (!q_xy || !msg_off || !rs_be || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    if (n == 0) return ECB_OK;
    for (size_t i = 0; i < n; i++)
        if (msg_off[i + 1] < msg_off[i]) return set_err(ctx, ECB_ERR_INVALID_ARG, "message offsets must be non-decreasing");
    if (msg_off[n] > msg_off[0] && !msgs) return set_err(ctx, ECB_ERR_INVALID_ARG, "null message buffer");
    int nd = (int)ctx->devs.size();
    std::vector<int> rc(nd, ECB_OK);
    std::vector<unsigned long long> bad(nd, ~0ull);
    auto worker = [&](int di) {
        DevCtx& d = *ctx->devs[di];
        size_t lo = n * (size_t)di / nd, hi = n * (size_t)(di + 1) / nd;
        if (lo == hi) return;
        std::lock_guard<std::mutex> g(d.mu);
        auto retire = [&](Slot& sl) -> int {
            if (!sl.busy) return ECB_OK;
            sl.busy = false;
            CU(cudaStreamSynchronize(sl.stream));
            if (*sl.h_status != ~0ull) {
                unsigned long long v = *sl.h_status, w = (((v >> 8) + sl.c0) << 8) | (v & 0xff);
                if (w < bad[di]) bad[di] = w;
            }
            return ECB_OK;
        };
        auto body = [&]() -> int {
            CU(cudaSetDevice(d.dev));
            SlotDrain drain{d};
            size_t ci = 0;
            for (size_t c0 = lo; c0 < hi && bad[di] == ~0ull; c0 += ctx->opt_chunk, ci++) {
                size_t cn = hi - c0 < ctx->opt_chunk ? hi - c0 : ctx->opt_chunk;
                Slot& sl = d.slots[ci % ECB_NSLOT];
                TRY(retire(sl));
                d.cur = &sl;
                size_t mbytes = (size_t)(msg_off[c0 + cn] - msg_off[c0]);
                TRY(ensure(ctx, sl.in[0], cn * 2 * fb));
                TRY(ensure(ctx, sl.in[1], cn * 2 * sb));
                TRY(ensure(ctx, sl.in[2], (cn + 1) * sizeof(uint64_t)));
                TRY(ensure(ctx, sl.out[1], mbytes + 16));
                TRY(ensure(ctx, sl.out[0], cn));
                CU(cudaMemcpyAsync(sl.in[0].p, q_xy + c0 * 2 * fb, cn * 2 * fb, cudaMemcpyHostToDevice, sl.stream));
                CU(cudaMemcpyAsync(sl.in[1].p, rs_be + c0 * 2 * sb, cn * 2 * sb, cudaMemcpyHostToDevice, sl.stream));
                CU(cudaMemcpyAsync(sl.in[2].p, msg_off + c0, (cn + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, sl.stream));
                if (mbytes) CU(cudaMemcpyAsync(sl.out[1].p, msgs + msg_off[c0], mbytes, cudaMemcpyHostToDevice, sl.stream));
                const unsigned char* d_msgs = (const unsigned char*)sl.out[1].p - msg_off[c0];
                if (curve == ECB_CURVE_P256R1)
                    TRY(dev_ecdsa_msgs_p256(ctx, d, (const u32*)sl.in[0].p, d_msgs, (const unsigned long long*)sl.in[2].p, hash,
                                            (const u32*)sl.in[1].p, cn, (unsigned char*)sl.out[0].p, sl.stream));
                else
                    TRY(dev_ecdsa_msgs_p384(ctx, d, (const u32*)sl.in[0].p, d_msgs, (const unsigned long long*)sl.in[2].p, hash,
                                            (const u32*)sl.in[1].p, cn, (unsigned char*)sl.out[0].p, sl.stream));
                CU(cudaMemcpyAsync(ok + c0, sl.out[0].p, cn, cudaMemcpyDeviceToHost, sl.stream));
                CU(cudaMemcpyAsync(sl.h_status, sl.d_status, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sl.stream));
                sl.busy = true;
                sl.c0 = c0;
            }
            int r2 = ECB_OK;
            for (Slot& sl : d.slots) {
                int r3 = retire(sl);
                if (r2 == ECB_OK) r2 = r3;
            }
            d.cur = &d.slots[0];
            drain.armed = r2 != ECB_OK;
            return r2;
        };
        rc[di] = body();
    };
    if (nd == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nd; i++) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < nd; i++)
        if (rc[i] != ECB_OK) return rc[i];
    unsigned long long first = ~0ull;
    for (int i = 0; i < nd; i++)
        if (bad[i] < first) first = bad[i];
    if (first != ~0ull) {
        if (bad_index) *bad_index = (size_t)(first >> 8);
        return set_err(ctx, ECB_ERR_POINT_NOT_ON_CURVE, "public key not on curve");
    }
    return ECB_OK;
}
int ecb_x25519_base(ecb_ctx* ctx, const uint8_t* k, size_t n, uint8_t* out) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!k || !out)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k, 32}}, {{out, 32}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_x25519_base(ctx, d, (const u32*)in[0], cn, (u32*)o[0], s);
                       });
}
int ecb_x448(ecb_ctx* ctx, const uint8_t* k, const uint8_t* u, size_t n, uint8_t* out) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!k || !u || !out)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k, 56}, {u, 56}}, {{out, 56}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_x448(ctx, d, (const u32*)in[0], (const u32*)in[1], cn, (u32*)o[0], s);
                       });
}
int ecb_wei_mul(ecb_ctx* ctx, int curve, const uint8_t* k_be, const uint8_t* xy_be, const uint8_t* inf_in, size_t n,
                uint8_t* out_xy, uint8_t* out_inf, size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    size_t fb, sb;
    if (curve_sizes(curve, fb, sb)) return set_err(ctx, ECB_ERR_INVALID_ARG, "unknown curve id");
    if (n && (!k_be || !xy_be || !out_xy)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k_be, sb}, {xy_be, 2 * fb}, {inf_in, 1}}, {{out_xy, 2 * fb}, {out_inf, 1}}, true, bad_index,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_wei_mul(ctx, d, curve, (const u32*)in[0], (const u32*)in[1], (const unsigned char*)in[2],
                                              cn, (u32*)o[0], (unsigned char*)o[1], s);
                       });
}
int ecb_wei_mul_base(ecb_ctx* ctx, int curve, const uint8_t* k_be, size_t n, uint8_t* out_xy, uint8_t* out_inf,
                     size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    size_t fb, sb;
    if (curve_sizes(curve, fb, sb)) return set_err(ctx, ECB_ERR_INVALID_ARG, "unknown curve id");
    if (n && (!k_be || !out_xy)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k_be, sb}}, {{out_xy, 2 * fb}, {out_inf, 1}}, true, bad_index,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_wei_mul_base(ctx, d, curve, (const u32*)in[0], cn, (u32*)o[0], (unsigned char*)o[1], s);
                       });
}
static int ecdsa_sign_hashed_impl(ecb_ctx* ctx, int curve, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* z_be, size_t n, uint8_t* rs_be,
                                  uint8_t* ok, bool ct);
int ecb_ecdsa_sign_hashed(ecb_ctx* ctx, int curve, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* z_be, size_t n, uint8_t* rs_be,
                          uint8_t* ok) {
    return ecdsa_sign_hashed_impl(ctx, curve, d_be, k_be, z_be, n, rs_be, ok, true);
}
int ecb_ecdsa_sign_hashed_vartime(ecb_ctx* ctx, int curve, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* z_be, size_t n, uint8_t* rs_be,
                                  uint8_t* ok) {
    return ecdsa_sign_hashed_impl(ctx, curve, d_be, k_be, z_be, n, rs_be, ok, false);
}
static int ecdsa_sign_hashed_impl(ecb_ctx* ctx, int curve, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* z_be, size_t n, uint8_t* rs_be,
                                  uint8_t* ok, bool ct) {
    if (!ctx) return ECB_ERR_CUDA;
    size_t fb, sb;
    if (curve_sizes(curve, fb, sb) || (curve != ECB_CURVE_P256R1 && curve != ECB_CURVE_P384R1)) return set_err(ctx, ECB_ERR_INVALID_ARG, "ECDSA is defined for p256r1/p384r1 only");
    if (n && (!d_be || !k_be || !z_be || !rs_be || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{d_be, sb}, {k_be, sb}, {z_be, sb}}, {{rs_be, 2 * sb}, {ok, 1}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           if (curve == ECB_CURVE_P256R1)
                               return dev_ecdsa_sign_p256(ctx, d, (const u32*)in[0], (const u32*)in[1], (const u32*)in[2], cn, (u32*)o[0],
                                                          (unsigned char*)o[1], s, ct);
                           return dev_ecdsa_sign_p384(ctx, d, (const u32*)in[0], (const u32*)in[1], (const u32*)in[2], cn, (u32*)o[0],
                                                      (unsigned char*)o[1], s, ct);
                       });
}
// ecdsa::sign on raw messages (ragged input): the chunk loop of ecb_ecdsa_verify; per chunk the slot holds
// d in in[0], k in in[1], the offsets in in[2], z (device-made) in in[3], the message bytes in `scratch`.
static int ecdsa_sign_impl(ecb_ctx* ctx, int curve, int hash, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* msgs, const uint64_t* msg_off,
                           size_t n, uint8_t* rs_be, uint8_t* ok, bool ct);
int ecb_ecdsa_sign(ecb_ctx* ctx, int curve, int hash, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* msgs, const uint64_t* msg_off,
                   size_t n, uint8_t* rs_be, uint8_t* ok) {
    return ecdsa_sign_impl(ctx, curve, hash, d_be, k_be, msgs, msg_off, n, rs_be, ok, true);
}
int ecb_ecdsa_sign_vartime(ecb_ctx* ctx, int curve, int hash, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* msgs, const uint64_t* msg_off,
                           size_t n, uint8_t* rs_be, uint8_t* ok) {
    return ecdsa_sign_impl(ctx, curve, hash, d_be, k_be, msgs, msg_off, n, rs_be, ok, false);
}
static int ecdsa_sign_impl(ecb_ctx* ctx, int curve, int hash, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* msgs, const uint64_t* msg_off,
                           size_t n, uint8_t* rs_be, uint8_t* ok, bool ct) {
    if (!ctx) return ECB_ERR_CUDA;
    size_t fb, sb;
    if (curve_sizes(curve, fb, sb) || (curve != ECB_CURVE_P256R1 && curve != ECB_CURVE_P384R1)) return set_err(ctx, ECB_ERR_INVALID_ARG, "ECDSA is defined for p256r1/p384r1 only");
    if (hash != 256 && hash != 384 && hash != 512) return set_err(ctx, ECB_ERR_INVALID_ARG, "hash must be 256, 384 or 512");
    if (n && (!d_be || !k_be || !msg_off || !rs_be || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    if (n == 0) return ECB_OK;
    for (size_t i = 0; i < n; i++)
        if (msg_off[i + 1] < msg_off[i]) return set_err(ctx, ECB_ERR_INVALID_ARG, "message offsets must be non-decreasing");
    if (msg_off[n] > msg_off[0] && !msgs) return set_err(ctx, ECB_ERR_INVALID_ARG, "null message buffer");
    int nd = (int)ctx->devs.size();
    std::vector<int> rc(nd, ECB_OK);
    auto worker = [&](int di) {
        DevCtx& d = *ctx->devs[di];
        size_t lo = n * (size_t)di / nd, hi = n * (size_t)(di + 1) / nd;
        if (lo == hi) return;
        std::lock_guard<std::mutex> g(d.mu);
        auto body = [&]() -> int {
            CU(cudaSetDevice(d.dev));
            SlotDrain drain{d};
            size_t ci = 0;
            for (size_t c0 = lo; c0 < hi; c0 += ctx->opt_chunk, ci++) {
                size_t cn = hi - c0 < ctx->opt_chunk ? hi - c0 : ctx->opt_chunk;
                Slot& sl = d.slots[ci % ECB_NSLOT];
                if (sl.busy) {
                    sl.busy = false;
                    CU(cudaStreamSynchronize(sl.stream));
                }
                d.cur = &sl;
                size_t mbytes = (size_t)(msg_off[c0 + cn] - msg_off[c0]);
                TRY(ensure(ctx, sl.in[0], cn * sb));
                TRY(ensure(ctx, sl.in[1], cn * sb));
                TRY(ensure(ctx, sl.in[2], (cn + 1) * sizeof(uint64_t)));
                TRY(ensure(ctx, sl.scratch, mbytes + 16));
                TRY(ensure(ctx, sl.out[0], cn * 2 * sb));
                TRY(ensure(ctx, sl.out[1], cn));
                CU(cudaMemcpyAsync(sl.in[0].p, d_be + c0 * sb, cn * sb, cudaMemcpyHostToDevice, sl.stream));
                CU(cudaMemcpyAsync(sl.in[1].p, k_be + c0 * sb, cn * sb, cudaMemcpyHostToDevice, sl.stream));
                CU(cudaMemcpyAsync(sl.in[2].p, msg_off + c0, (cn + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, sl.stream));
                if (mbytes) CU(cudaMemcpyAsync(sl.scratch.p, msgs + msg_off[c0], mbytes, cudaMemcpyHostToDevice, sl.stream));
                const unsigned char* d_msgs = (const unsigned char*)sl.scratch.p - msg_off[c0];
                if (curve == ECB_CURVE_P256R1)
                    TRY(dev_ecdsa_sign_msgs_p256(ctx, d, (const u32*)sl.in[0].p, (const u32*)sl.in[1].p, d_msgs, (const unsigned long long*)sl.in[2].p,
                                                 hash, cn, (u32*)sl.out[0].p, (unsigned char*)sl.out[1].p, sl.stream, ct));
                else
                    TRY(dev_ecdsa_sign_msgs_p384(ctx, d, (const u32*)sl.in[0].p, (const u32*)sl.in[1].p, d_msgs, (const unsigned long long*)sl.in[2].p,
                                                 hash, cn, (u32*)sl.out[0].p, (unsigned char*)sl.out[1].p, sl.stream, ct));
                CU(cudaMemcpyAsync(rs_be + c0 * 2 * sb, sl.out[0].p, cn * 2 * sb, cudaMemcpyDeviceToHost, sl.stream));
                CU(cudaMemcpyAsync(ok + c0, sl.out[1].p, cn, cudaMemcpyDeviceToHost, sl.stream));
                sl.busy = true;
            }
            for (Slot& sl : d.slots) {
                if (!sl.busy) continue;
                sl.busy = false;
                CU(cudaStreamSynchronize(sl.stream));
            }
            d.cur = &d.slots[0];
            drain.armed = false;
            return ECB_OK;
        };
        rc[di] = body();
    };
    if (nd == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nd; i++) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < nd; i++)
        if (rc[i] != ECB_OK) return rc[i];
    return ECB_OK;
}
int ecb_wei_decompress(ecb_ctx* ctx, int curve, const uint8_t* x_be, const uint8_t* sign, size_t n, uint8_t* out_xy, uint8_t* ok) {
    if (!ctx) return ECB_ERR_CUDA;
    size_t fb, sb;
    if (curve_sizes(curve, fb, sb)) return set_err(ctx, ECB_ERR_INVALID_ARG, "unknown curve id");
    if (n && (!x_be || !sign || !out_xy || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{x_be, fb}, {sign, 1}}, {{out_xy, 2 * fb}, {ok, 1}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           switch (curve) {
                               case ECB_CURVE_P256R1:
                                   return dev_wei_decompress_p256(ctx, d, (const u32*)in[0], (const unsigned char*)in[1], cn, (u32*)o[0],
                                                                  (unsigned char*)o[1], s);
                               case ECB_CURVE_P384R1:
                                   return dev_wei_decompress_p384(ctx, d, (const u32*)in[0], (const unsigned char*)in[1], cn, (u32*)o[0],
                                                                  (unsigned char*)o[1], s);
                               case ECB_CURVE_P256K1:
                                   return dev_wei_decompress_k256(ctx, d, (const u32*)in[0], (const unsigned char*)in[1], cn, (u32*)o[0],
                                                                  (unsigned char*)o[1], s);
                               default:
                                   return dev_wei_decompress_bls(ctx, d, (const u32*)in[0], (const unsigned char*)in[1], cn, (u32*)o[0],
                                                                 (unsigned char*)o[1], s);
                           }
                       });
}
int ecb_bls12_381_g1_from_compressed(ecb_ctx* ctx, const uint8_t* enc, size_t n, int check_subgroup, uint8_t* out_xy, uint8_t* ok) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!enc || !out_xy || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{enc, 48}}, {{out_xy, 96}, {ok, 1}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_bls_g1_from_compressed(ctx, d, (const u32*)in[0], cn, check_subgroup ? 1 : 0, (u32*)o[0],
                                                             (unsigned char*)o[1], s);
                       });
}
int ecb_bls12_381_g1_from_uncompressed(ecb_ctx* ctx, const uint8_t* enc, size_t n, int check_subgroup, uint8_t* out_xy, uint8_t* out_inf,
                                       uint8_t* ok) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!enc || !out_xy || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{enc, 96}}, {{out_xy, 96}, {ok, 1}, {out_inf, 1}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_bls_g1_from_uncompressed(ctx, d, (const u32*)in[0], cn, check_subgroup ? 1 : 0, (u32*)o[0],
                                                               (unsigned char*)o[2], (unsigned char*)o[1], s);
                       });
}
int ecb_bls12_381_g1_to_uncompressed(ecb_ctx* ctx, const uint8_t* xy_be, const uint8_t* inf, size_t n, uint8_t* enc) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!xy_be || !enc)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{xy_be, 96}, {inf, 1}}, {{enc, 96}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_bls_g1_to_uncompressed(ctx, d, (const u32*)in[0], (const unsigned char*)in[1], cn, (u32*)o[0], s);
                       });
}
// Multi-scalar multiplication sum_i k_i P_i (csrc/msm.cuh): every device reduces its contiguous slice to one Jacobian
// point, the partial sums are copied to the first device (the one cross-device step of this library) and added there.
int ecb_wei_msm(ecb_ctx* ctx, int curve, const uint8_t* k_be, const uint8_t* xy_be, size_t n, uint8_t* out_xy, uint8_t* out_inf,
                size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    if (bad_index) *bad_index = (size_t)-1;
    size_t fb, sb;
    if (curve_sizes(curve, fb, sb) || (curve != ECB_CURVE_BLS12_381_G1 && curve != ECB_CURVE_P256K1))
        return set_err(ctx, ECB_ERR_INVALID_ARG, "ecb_wei_msm: bls12_381_g1 and p256k1 only");
    if (!out_xy || !out_inf || (n && (!k_be || !xy_be))) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    if (n >= ((size_t)1 << 31)) return set_err(ctx, ECB_ERR_INVALID_ARG, "ecb_wei_msm: n must be below 2^31");
    if (n == 0) {   // the empty sum is the identity
        memset(out_xy, 0, 2 * fb);
        *out_inf = 1;
        return ECB_OK;
    }
    const int nd_all = (int)ctx->devs.size();
    const int nd = n < (size_t)nd_all * 1024 ? 1 : nd_all;   // tiny batches stay on one device
    const size_t pw = 3 * (fb / 4) * sizeof(u32);
    std::vector<int> rc(nd, ECB_OK);
    std::vector<unsigned long long> bad(nd, ~0ull);
    DevCtx& d0 = *ctx->devs[0];
    {   // room for the partial sums on the first device
        std::lock_guard<std::mutex> g(d0.mu);
        if (cudaSetDevice(d0.dev) != cudaSuccess) return set_err(ctx, ECB_ERR_CUDA, "cudaSetDevice");
        d0.cur = &d0.slots[0];
        TRY(ensure(ctx, d0.slots[0].in[2], (size_t)nd * pw));
    }
    auto worker = [&](int di) {
        DevCtx& d = *ctx->devs[di];
        size_t lo = n * (size_t)di / nd, hi = n * (size_t)(di + 1) / nd, cn = hi - lo;
        std::lock_guard<std::mutex> g(d.mu);
        auto body = [&]() -> int {
            CU(cudaSetDevice(d.dev));
            Slot& sl = d.slots[0];
            d.cur = &sl;
            cudaStream_t s = sl.stream;
            TRY(ensure(ctx, sl.in[0], cn * sb));
            TRY(ensure(ctx, sl.in[1], cn * 2 * fb));
            TRY(ensure(ctx, sl.out[0], pw));
            CU(cudaMemcpyAsync(sl.in[0].p, k_be + lo * sb, cn * sb, cudaMemcpyHostToDevice, s));
            CU(cudaMemcpyAsync(sl.in[1].p, xy_be + lo * 2 * fb, cn * 2 * fb, cudaMemcpyHostToDevice, s));
            if (curve == ECB_CURVE_BLS12_381_G1) TRY(dev_wei_msm_bls(ctx, d, (const u32*)sl.in[0].p, (const u32*)sl.in[1].p, cn, (u32*)sl.out[0].p, s));
            else TRY(dev_wei_msm_k256(ctx, d, (const u32*)sl.in[0].p, (const u32*)sl.in[1].p, cn, (u32*)sl.out[0].p, s));
            CU(cudaMemcpyAsync(sl.h_status, sl.d_status, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
            CU(cudaMemcpyPeerAsync((char*)d0.slots[0].in[2].p + (size_t)di * pw, d0.dev, sl.out[0].p, d.dev, pw, s));
            CU(cudaStreamSynchronize(s));
            if (*sl.h_status != ~0ull) bad[di] = (((*sl.h_status >> 8) + lo) << 8) | (*sl.h_status & 0xff);
            return ECB_OK;
        };
        rc[di] = body();
        if (rc[di] != ECB_OK) cudaStreamSynchronize(d.slots[0].stream);
    };
    if (nd == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nd; i++) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
    }
    for (int i = 0; i < nd; i++)
        if (rc[i] != ECB_OK) return rc[i];
    unsigned long long first = ~0ull;
    for (int i = 0; i < nd; i++)
        if (bad[i] < first) first = bad[i];
    if (first != ~0ull) {
        if (bad_index) *bad_index = (size_t)(first >> 8);
        int code = (first & 0xff) == ecb::ST_NONCANONICAL_SCALAR ? ECB_ERR_NONCANONICAL_SCALAR : ECB_ERR_POINT_NOT_ON_CURVE;
        return set_err(ctx, code, code == ECB_ERR_NONCANONICAL_SCALAR ? "non-canonical scalar" : "point not on curve");
    }
    std::lock_guard<std::mutex> g(d0.mu);
    CU(cudaSetDevice(d0.dev));
    Slot& sl = d0.slots[0];
    d0.cur = &sl;
    TRY(ensure(ctx, sl.out[1], 2 * fb + 16));
    u32* o = (u32*)sl.out[1].p;
    unsigned char* oi = (unsigned char*)sl.out[1].p + 2 * fb;
    if (curve == ECB_CURVE_BLS12_381_G1) TRY(dev_wei_msm_finish_bls(ctx, d0, (const u32*)sl.in[2].p, nd, o, oi, sl.stream));
    else TRY(dev_wei_msm_finish_k256(ctx, d0, (const u32*)sl.in[2].p, nd, o, oi, sl.stream));
    CU(cudaMemcpyAsync(out_xy, o, 2 * fb, cudaMemcpyDeviceToHost, sl.stream));
    CU(cudaMemcpyAsync(out_inf, oi, 1, cudaMemcpyDeviceToHost, sl.stream));
    CU(cudaStreamSynchronize(sl.stream));
    if (sl.hi) CU(cudaStreamSynchronize(sl.hi));
    return ECB_OK;
}
// ristretto255 (src/curve/curve25519/ristretto255.rs): encodings and the scalar multiplications between them
int ecb_ristretto255_decompress(ecb_ctx* ctx, const uint8_t* enc, size_t n, uint8_t* xy_le, uint8_t* ok) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!enc || !xy_le || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{enc, 32}}, {{xy_le, 64}, {ok, 1}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_ristretto255_decompress(ctx, d, (const u32*)in[0], cn, (u32*)o[0], (unsigned char*)o[1], s);
                       });
}
int ecb_ristretto255_compress(ecb_ctx* ctx, const uint8_t* xy_le, size_t n, uint8_t* enc) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!xy_le || !enc)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{xy_le, 64}}, {{enc, 32}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_ristretto255_compress(ctx, d, (const u32*)in[0], cn, (u32*)o[0], s);
                       });
}
int ecb_ristretto255_mul(ecb_ctx* ctx, const uint8_t* k_le, const uint8_t* enc_in, size_t n, uint8_t* enc_out, size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!k_le || !enc_in || !enc_out)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k_le, 32}, {enc_in, 32}}, {{enc_out, 32}}, true, bad_index,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_ristretto255_mul(ctx, d, (const u32*)in[0], (const u32*)in[1], cn, (u32*)o[0], s);
                       });
}
int ecb_ristretto255_mul_base(ecb_ctx* ctx, const uint8_t* k_le, size_t n, uint8_t* enc_out, size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!k_le || !enc_out)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{k_le, 32}}, {{enc_out, 32}}, true, bad_index,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_ristretto255_mul_base(ctx, d, (const u32*)in[0], cn, (u32*)o[0], s);
                       });
}
// decode_point / Point::decompress over a batch (protocol/ed25519.rs:38-59, curve25519.rs:772)
int ecb_ed25519_decompress(ecb_ctx* ctx, const uint8_t* enc, size_t n, uint8_t* xy_le, uint8_t* ok) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!enc || !xy_le || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{enc, 32}}, {{xy_le, 64}, {ok, 1}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_ed25519_decompress(ctx, d, (const u32*)in[0], cn, (u32*)o[0], (unsigned char*)o[1], s);
                       });
}
int ecb_bls12_381_g1_to_compressed(ecb_ctx* ctx, const uint8_t* xy_be, const uint8_t* inf, size_t n, uint8_t* enc) {
    if (!ctx) return ECB_ERR_CUDA;
    if (n && (!xy_be || !enc)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{xy_be, 96}, {inf, 1}}, {{enc, 48}}, false, nullptr,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           return dev_bls_g1_to_compressed(ctx, d, (const u32*)in[0], (const unsigned char*)in[1], cn, (u32*)o[0], s);
                       });
}
int ecb_ecdsa_verify_hashed(ecb_ctx* ctx, int curve, const uint8_t* q_xy, const uint8_t* z_be, const uint8_t* rs_be,
                            size_t n, uint8_t* ok, size_t* bad_index) {
    if (!ctx) return ECB_ERR_CUDA;
    size_t fb, sb;
    if (curve_sizes(curve, fb, sb)) return set_err(ctx, ECB_ERR_INVALID_ARG, "unknown curve id");
    if (curve != ECB_CURVE_P256R1 && curve != ECB_CURVE_P384R1) return set_err(ctx, ECB_ERR_INVALID_ARG, "ECDSA is defined for p256r1/p384r1 only");
    if (n && (!q_xy || !z_be || !rs_be || !ok)) return set_err(ctx, ECB_ERR_INVALID_ARG, "null buffer");
    return run_sharded(ctx, n, {{q_xy, 2 * fb}, {z_be, sb}, {rs_be, 2 * sb}}, {{ok, 1}}, true, bad_index,
                       [&](DevCtx& d, cudaStream_t s, const void** in, void** o, size_t cn) {
                           if (curve == ECB_CURVE_P256R1)
                               return dev_ecdsa_p256(ctx, d, (const u32*)in[0], (const u32*)in[1], (const u32*)in[2], cn,
                                                             (unsigned char*)o[0], s);
                           return dev_ecdsa_p384(ctx, d, (const u32*)in[0], (const u32*)in[1], (const u32*)in[2], cn,
                                                         (unsigned char*)o[0], s);
                       });
}

// ---- device-resident variants ---------------------------------------------------------------
static inline void single_slot(DevCtx* d) {  // *_dev calls that do not fork: only slot 0's status word is live
    d->cur = &d->slots[0];
    d->slots[0].c0 = 0;
    d->dev_slots_used = 1;
}
static DevCtx* get_dev(ecb_ctx* ctx, int i) { return (ctx && i >= 0 && i < (int)ctx->devs.size()) ? ctx->devs[i] : nullptr; }
// *_dev entry points only enqueue: while one runs, ensure() and the table builders refuse to allocate
// (ECB_ERR_NOT_READY) unless ecb_warm() is the caller.  Successive *_dev calls share the work buffers of
// slot 0: enqueue them on ONE stream, or order streams with events.
struct DevCallScope {
    ecb_ctx* ctx;
    bool prev;
    explicit DevCallScope(ecb_ctx* c, bool warming) : ctx(c), prev(c->no_alloc) { ctx->no_alloc = !warming; }
    ~DevCallScope() { ctx->no_alloc = prev; }
};
static thread_local bool g_warming = false;
#define DEV_ENTER(n_)                          \
    if (!d) return ECB_ERR_INVALID_ARG;        \
    if ((n_) == 0) return ECB_OK;              \
    CU(cudaSetDevice(d->dev));                 \
    DevCallScope scope_(ctx, g_warming)

int ecb_ed25519_mul_base_dev(ecb_ctx* ctx, int di, const void* d_k, size_t n, void* d_xy, void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    return dev_forkjoin(ctx, *d, n, (cudaStream_t)stream, [&](size_t lo, size_t cnt, cudaStream_t st) {
        return dev_ed25519_mul_base(ctx, *d, (const u32*)d_k + lo * 8, cnt, (u32*)d_xy + lo * 16, false, st);
    });
}
int ecb_ed25519_mul_dev(ecb_ctx* ctx, int di, const void* d_k, const void* d_xy_in, size_t n, void* d_xy_out, void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    return dev_ed25519_mul(ctx, *d, (const u32*)d_k, (const u32*)d_xy_in, n, (u32*)d_xy_out, (cudaStream_t)stream);
}
int ecb_x25519_dev(ecb_ctx* ctx, int di, const void* d_k, const void* d_u, size_t n, void* d_out, void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    return dev_x25519(ctx, *d, (const u32*)d_k, (const u32*)d_u, n, (u32*)d_out, (cudaStream_t)stream);
}
int ecb_x25519_base_dev(ecb_ctx* ctx, int di, const void* d_k, size_t n, void* d_out, void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    return dev_x25519_base(ctx, *d, (const u32*)d_k, n, (u32*)d_out, (cudaStream_t)stream);
}
int ecb_wei_mul_dev(ecb_ctx* ctx, int di, int curve, const void* d_k, const void* d_xy, size_t n, void* d_out, void* d_inf,
                    void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    return dev_wei_mul(ctx, *d, curve, (const u32*)d_k, (const u32*)d_xy, nullptr, n, (u32*)d_out, (unsigned char*)d_inf,
                       (cudaStream_t)stream);
}
int ecb_wei_mul_base_dev(ecb_ctx* ctx, int di, int curve, const void* d_k, size_t n, void* d_out, void* d_inf, void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    size_t fb, sb;
    if (curve_sizes(curve, fb, sb)) return set_err(ctx, ECB_ERR_INVALID_ARG, "unknown curve id");
    return dev_forkjoin(ctx, *d, n, (cudaStream_t)stream, [&](size_t lo, size_t cnt, cudaStream_t st) {
        return dev_wei_mul_base(ctx, *d, curve, (const u32*)d_k + lo * (sb / 4), cnt, (u32*)d_out + lo * (fb / 2),
                                d_inf ? (unsigned char*)d_inf + lo : nullptr, st);
    });
}
int ecb_wei_decompress_dev(ecb_ctx* ctx, int di, int curve, const void* d_x, const void* d_sign, size_t n, void* d_out, void* d_ok,
                           void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    switch (curve) {
        case ECB_CURVE_P256R1:
            return dev_wei_decompress_p256(ctx, *d, (const u32*)d_x, (const unsigned char*)d_sign, n, (u32*)d_out, (unsigned char*)d_ok,
                                           (cudaStream_t)stream);
        case ECB_CURVE_P384R1:
            return dev_wei_decompress_p384(ctx, *d, (const u32*)d_x, (const unsigned char*)d_sign, n, (u32*)d_out, (unsigned char*)d_ok,
                                           (cudaStream_t)stream);
        case ECB_CURVE_BLS12_381_G1:
            return dev_wei_decompress_bls(ctx, *d, (const u32*)d_x, (const unsigned char*)d_sign, n, (u32*)d_out, (unsigned char*)d_ok,
                                          (cudaStream_t)stream);
        case ECB_CURVE_P256K1:
            return dev_wei_decompress_k256(ctx, *d, (const u32*)d_x, (const unsigned char*)d_sign, n, (u32*)d_out, (unsigned char*)d_ok,
                                           (cudaStream_t)stream);
    }
    return set_err(ctx, ECB_ERR_INVALID_ARG, "unknown curve id");
}
int ecb_bls12_381_g1_from_compressed_dev(ecb_ctx* ctx, int di, const void* d_enc, size_t n, int check_subgroup, void* d_out, void* d_ok,
                                         void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    return dev_bls_g1_from_compressed(ctx, *d, (const u32*)d_enc, n, check_subgroup ? 1 : 0, (u32*)d_out, (unsigned char*)d_ok,
                                      (cudaStream_t)stream);
}
static int ecdsa_sign_hashed_dev_impl(ecb_ctx* ctx, int di, int curve, const void* d_d, const void* d_k, const void* d_z, size_t n, void* d_rs,
                                      void* d_ok, void* stream, bool ct) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    if (curve == ECB_CURVE_P256R1)
        return dev_ecdsa_sign_p256(ctx, *d, (const u32*)d_d, (const u32*)d_k, (const u32*)d_z, n, (u32*)d_rs, (unsigned char*)d_ok,
                                   (cudaStream_t)stream, ct);
    if (curve == ECB_CURVE_P384R1)
        return dev_ecdsa_sign_p384(ctx, *d, (const u32*)d_d, (const u32*)d_k, (const u32*)d_z, n, (u32*)d_rs, (unsigned char*)d_ok,
                                   (cudaStream_t)stream, ct);
    return set_err(ctx, ECB_ERR_INVALID_ARG, "ECDSA is defined for p256r1/p384r1 only");
}
int ecb_ecdsa_sign_hashed_dev(ecb_ctx* ctx, int di, int curve, const void* d_d, const void* d_k, const void* d_z, size_t n, void* d_rs,
                              void* d_ok, void* stream) {
    return ecdsa_sign_hashed_dev_impl(ctx, di, curve, d_d, d_k, d_z, n, d_rs, d_ok, stream, true);
}
int ecb_ecdsa_sign_hashed_vartime_dev(ecb_ctx* ctx, int di, int curve, const void* d_d, const void* d_k, const void* d_z, size_t n, void* d_rs,
                                      void* d_ok, void* stream) {
    return ecdsa_sign_hashed_dev_impl(ctx, di, curve, d_d, d_k, d_z, n, d_rs, d_ok, stream, false);
}
static int ed25519_public_from_seed_dev_impl(ecb_ctx* ctx, int di, const void* d_seeds, size_t n, void* d_pub, void* stream, bool ct) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    return dev_ed25519_public_from_seed(ctx, *d, (const unsigned char*)d_seeds, n, (u32*)d_pub, (cudaStream_t)stream, ct);
}
int ecb_ed25519_public_from_seed_dev(ecb_ctx* ctx, int di, const void* d_seeds, size_t n, void* d_pub, void* stream) {
    return ed25519_public_from_seed_dev_impl(ctx, di, d_seeds, n, d_pub, stream, true);
}
int ecb_ed25519_public_from_seed_vartime_dev(ecb_ctx* ctx, int di, const void* d_seeds, size_t n, void* d_pub, void* stream) {
    return ed25519_public_from_seed_dev_impl(ctx, di, d_seeds, n, d_pub, stream, false);
}
static int ed25519_sign_dev_impl(ecb_ctx* ctx, int di, const void* d_seeds, const void* d_pub, const void* d_msgs, const void* d_msg_off, size_t n,
                                 void* d_sig, void* stream, bool ct) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    return dev_ed25519_sign(ctx, *d, (const unsigned char*)d_seeds, (const unsigned char*)d_pub, (const unsigned char*)d_msgs,
                            (const unsigned long long*)d_msg_off, n, (unsigned char*)d_sig, (cudaStream_t)stream, ct);
}
int ecb_ed25519_sign_dev(ecb_ctx* ctx, int di, const void* d_seeds, const void* d_pub, const void* d_msgs, const void* d_msg_off, size_t n,
                         void* d_sig, void* stream) {
    return ed25519_sign_dev_impl(ctx, di, d_seeds, d_pub, d_msgs, d_msg_off, n, d_sig, stream, true);
}
int ecb_ed25519_sign_vartime_dev(ecb_ctx* ctx, int di, const void* d_seeds, const void* d_pub, const void* d_msgs, const void* d_msg_off,
                                 size_t n, void* d_sig, void* stream) {
    return ed25519_sign_dev_impl(ctx, di, d_seeds, d_pub, d_msgs, d_msg_off, n, d_sig, stream, false);
}
int ecb_x448_dev(ecb_ctx* ctx, int di, const void* d_k, const void* d_u, size_t n, void* d_out, void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    return dev_x448(ctx, *d, (const u32*)d_k, (const u32*)d_u, n, (u32*)d_out, (cudaStream_t)stream);
}
int ecb_ed25519_verify_prehashed_dev(ecb_ctx* ctx, int di, const void* d_a, const void* d_r, const void* d_s, const void* d_k,
                                     size_t n, void* d_ok, void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    return dev_ed25519_verify(ctx, *d, (const u32*)d_a, (const u32*)d_r, (const u32*)d_s, (const u32*)d_k, n,
                              (unsigned char*)d_ok, (cudaStream_t)stream);
}
int ecb_ecdsa_verify_hashed_dev(ecb_ctx* ctx, int di, int curve, const void* d_q, const void* d_z, const void* d_rs, size_t n,
                                void* d_ok, void* stream) {
    DevCtx* d = get_dev(ctx, di);
    DEV_ENTER(n);
    single_slot(d);
    if (curve == ECB_CURVE_P256R1)
        return dev_ecdsa_p256(ctx, *d, (const u32*)d_q, (const u32*)d_z, (const u32*)d_rs, n, (unsigned char*)d_ok, (cudaStream_t)stream);
    if (curve == ECB_CURVE_P384R1)
        return dev_ecdsa_p384(ctx, *d, (const u32*)d_q, (const u32*)d_z, (const u32*)d_rs, n, (unsigned char*)d_ok, (cudaStream_t)stream);
    return set_err(ctx, ECB_ERR_INVALID_ARG, "ECDSA is defined for p256r1/p384r1 only");
}
int ecb_profile_collect(ecb_ctx* ctx, int di, double* main_ms, double* fin_ms, int* calls) {
    DevCtx* d = get_dev(ctx, di);
    if (!d) return ECB_ERR_INVALID_ARG;
    CU(cudaSetDevice(d->dev));
    double m = 0, f = 0;
    int c = 0;
    for (auto& r : d->prof) {
        float t1 = 0, t2 = 0;
        CU(cudaEventSynchronize(r.c));
        CU(cudaEventElapsedTime(&t1, r.a, r.b));
        CU(cudaEventElapsedTime(&t2, r.b, r.c));
        m += t1;
        f += t2;
        c++;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
        cudaEventDestroy(r.c);
    }
    d->prof.clear();
    if (main_ms) *main_ms = m;
    if (fin_ms) *fin_ms = f;
    if (calls) *calls = c;
    return ECB_OK;
}
int ecb_dev_status(ecb_ctx* ctx, int di, size_t* bad_index) {
    DevCtx* d = get_dev(ctx, di);
    if (!d) return ECB_ERR_INVALID_ARG;
    CU(cudaSetDevice(d->dev));
    unsigned long long v = ~0ull;
    for (int k = 0; k < d->dev_slots_used; k++) {
        Slot& sl = d->slots[k];
        CU(cudaMemcpy(sl.h_status, sl.d_status, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (*sl.h_status == ~0ull) continue;
        unsigned long long w = (((*sl.h_status >> 8) + sl.c0) << 8) | (*sl.h_status & 0xff);
        if (w < v) v = w;
    }
    if (bad_index) *bad_index = (size_t)-1;
    if (v == ~0ull) return ECB_OK;
    if (bad_index) *bad_index = (size_t)(v >> 8);
    return (v & 0xff) == ecb::ST_NONCANONICAL_SCALAR ? ECB_ERR_NONCANONICAL_SCALAR : ECB_ERR_POINT_NOT_ON_CURVE;
}

int ecb_imad_probe(ecb_ctx* ctx, int di, int variant, int iters, double* macs_per_s, double* ms_out) {
    DevCtx* d = get_dev(ctx, di);
    if (!d) return ECB_ERR_INVALID_ARG;
    CU(cudaSetDevice(d->dev));
    return dev_imad_probe(ctx, *d, variant, iters, macs_per_s, ms_out);
}

int ecb_latency_probe(ecb_ctx* ctx, int di, int variant, int threads, int reps, double* cycles, double* sm_mhz) {
    DevCtx* d = get_dev(ctx, di);
    if (!d) return ECB_ERR_INVALID_ARG;
    CU(cudaSetDevice(d->dev));
    return dev_latency_probe(ctx, *d, variant, threads, reps, cycles, sm_mhz);
}

int ecb_fieldmul_probe(ecb_ctx* ctx, int di, int fp64_num, int fp64_den, int blocks_per_sm, int reps, double* muls_per_s, double* check) {
    DevCtx* d = get_dev(ctx, di);
    if (!d) return ECB_ERR_INVALID_ARG;
    CU(cudaSetDevice(d->dev));
    return dev_fieldmul_probe(ctx, *d, fp64_num, fp64_den, blocks_per_sm, reps, muls_per_s, check);
}

long ecb_debug_fused_trace(ecb_ctx* ctx, int di, unsigned long long* out, size_t cap_blocks) {
    DevCtx* d = get_dev(ctx, di);
    if (!d) return ECB_ERR_INVALID_ARG;
    if (cudaSetDevice(d->dev) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return ECB_ERR_CUDA;
    size_t nb = d->trace_blocks < cap_blocks ? d->trace_blocks : cap_blocks;
    if (out && nb && cudaMemcpy(out, d->trace.p, nb * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return ECB_ERR_CUDA;
    return (long)d->trace_blocks;
}

long ecb_debug_chunk_plan(size_t lo, size_t hi, size_t chunk, long ramp, size_t* bounds, size_t cap) {
    if (hi < lo || chunk == 0 || ramp < 0 || ramp > 4) return -1;
    std::vector<size_t> b;
    chunk_plan(lo, hi, chunk, ramp, b);
    if (bounds)
        for (size_t i = 0; i < b.size() && i < cap; i++) bounds[i] = b[i];
    return (long)b.size();
}
long ecb_debug_ed25519_table(ecb_ctx* ctx, int di, uint8_t* out, size_t cap, int* w, int* nwin) {
    DevCtx* d = get_dev(ctx, di);
    if (!d) return ECB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> g(d->mu);
    if (cudaSetDevice(d->dev) != cudaSuccess) return ECB_ERR_CUDA;
    if (!d->ed_table || (ctx->opt_ed_w && d->ed_w != (int)ctx->opt_ed_w) || d->ed_stride != (int)ctx->opt_ed_stride) {
        int r = dev_ed25519_build_table(ctx, *d, ctx->opt_ed_w ? (int)ctx->opt_ed_w : 16);
        if (r != ECB_OK) return r;
    }
    size_t ntab = (size_t)d->ed_nwin << (d->ed_w - 1);
    if (w) *w = d->ed_w;
    if (nwin) *nwin = d->ed_nwin;
    if (out) {   // always handed out packed (96 B per entry), whatever the device stride
        size_t cnt = cap / 96 < ntab ? cap / 96 : ntab;
        if (cudaMemcpy2D(out, 96, d->ed_table, (size_t)d->ed_stride * 4, 96, cnt, cudaMemcpyDeviceToHost) != cudaSuccess) return ECB_ERR_CUDA;
    }
    return (long)ntab;
}

// ---- explicit preparation -------------------------------------------------------------------------
// ecb_warm runs `op` once per device on an all-zero batch of max_n elements with allocation allowed:
// afterwards the generator comb the op needs is built and every work buffer is sized, so *_dev calls of
// that op with n <= max_n (same options) only enqueue kernels.
struct WarmOp { const char* name; size_t in_b[4]; size_t out_b[2]; };
static const WarmOp* warm_lookup(const char* op, int curve, WarmOp& tmp) {
    size_t fb = 0, sb = 0;
    const bool wei = curve_sizes(curve, fb, sb) == ECB_OK;
    static const WarmOp fixed[] = {
        {"ed25519_mul_base", {32, 0, 0, 0}, {64, 0}},       {"ed25519_mul", {32, 64, 0, 0}, {64, 0}},
        {"x25519", {32, 32, 0, 0}, {32, 0}},                {"x25519_base", {32, 0, 0, 0}, {32, 0}},
        {"x448", {56, 56, 0, 0}, {56, 0}},                  {"ed25519_verify_prehashed", {32, 32, 32, 32}, {1, 0}},
        {"ed25519_public_from_seed", {32, 0, 0, 0}, {32, 0}}, {"ed25519_sign", {32, 32, 64, 8}, {64, 0}},
        {"ed25519_public_from_seed_vartime", {32, 0, 0, 0}, {32, 0}}, {"ed25519_sign_vartime", {32, 32, 64, 8}, {64, 0}},
        {"bls12_381_g1_from_compressed", {48, 0, 0, 0}, {96, 1}},
    };
    for (const WarmOp& w : fixed)
        if (!strcmp(op, w.name)) return &w;
    if (!wei) return nullptr;
    tmp.name = op;
    if (!strcmp(op, "wei_mul")) tmp = WarmOp{op, {sb, 2 * fb, 0, 0}, {2 * fb, 1}};
    else if (!strcmp(op, "wei_mul_base")) tmp = WarmOp{op, {sb, 0, 0, 0}, {2 * fb, 1}};
    else if (!strcmp(op, "wei_decompress")) tmp = WarmOp{op, {fb, 1, 0, 0}, {2 * fb, 1}};
    else if (!strcmp(op, "ecdsa_verify_hashed")) tmp = WarmOp{op, {2 * fb, sb, 2 * sb, 0}, {1, 0}};
    else if (!strcmp(op, "ecdsa_sign_hashed") || !strcmp(op, "ecdsa_sign_hashed_vartime")) tmp = WarmOp{op, {sb, sb, sb, 0}, {2 * sb, 1}};
    else return nullptr;
    return &tmp;
}
int ecb_warm(ecb_ctx* ctx, const char* op, int curve, size_t max_n) {
    if (!ctx || !op) return ECB_ERR_INVALID_ARG;
    WarmOp tmp{};
    const WarmOp* w = warm_lookup(op, curve, tmp);
    if (!w) return set_err(ctx, ECB_ERR_INVALID_ARG, std::string("ecb_warm: unknown op ") + op);
    if (max_n == 0) max_n = 1;
    for (int di = 0; di < (int)ctx->devs.size(); di++) {
        DevCtx* d = ctx->devs[di];
        CU(cudaSetDevice(d->dev));
        void* in[4] = {nullptr, nullptr, nullptr, nullptr};
        void* out[2] = {nullptr, nullptr};
        int rc = ECB_OK;
        auto body = [&]() -> int {
            for (int k = 0; k < 4; k++)
                if (w->in_b[k]) {
                    size_t bytes = w->in_b[k] * (max_n + 1);
                    CU(cudaMalloc(&in[k], bytes));
                    CU(cudaMemset(in[k], 0, bytes));
                }
            for (int k = 0; k < 2; k++)
                if (w->out_b[k]) CU(cudaMalloc(&out[k], w->out_b[k] * max_n));
            void* st = (void*)d->stream;
            g_warming = true;
            int r;
            if (!strcmp(op, "ed25519_mul_base")) r = ecb_ed25519_mul_base_dev(ctx, di, in[0], max_n, out[0], st);
            else if (!strcmp(op, "ed25519_mul")) r = ecb_ed25519_mul_dev(ctx, di, in[0], in[1], max_n, out[0], st);
            else if (!strcmp(op, "x25519")) r = ecb_x25519_dev(ctx, di, in[0], in[1], max_n, out[0], st);
            else if (!strcmp(op, "x25519_base")) r = ecb_x25519_base_dev(ctx, di, in[0], max_n, out[0], st);
            else if (!strcmp(op, "x448")) r = ecb_x448_dev(ctx, di, in[0], in[1], max_n, out[0], st);
            else if (!strcmp(op, "ed25519_verify_prehashed")) r = ecb_ed25519_verify_prehashed_dev(ctx, di, in[0], in[1], in[2], in[3], max_n, out[0], st);
            else if (!strcmp(op, "ed25519_public_from_seed")) r = ecb_ed25519_public_from_seed_dev(ctx, di, in[0], max_n, out[0], st);
            else if (!strcmp(op, "ed25519_sign")) r = ecb_ed25519_sign_dev(ctx, di, in[0], in[1], in[2], in[3], max_n, out[0], st);   // all-zero offsets: empty messages
            else if (!strcmp(op, "ed25519_public_from_seed_vartime")) r = ecb_ed25519_public_from_seed_vartime_dev(ctx, di, in[0], max_n, out[0], st);
            else if (!strcmp(op, "ed25519_sign_vartime")) r = ecb_ed25519_sign_vartime_dev(ctx, di, in[0], in[1], in[2], in[3], max_n, out[0], st);
            else if (!strcmp(op, "bls12_381_g1_from_compressed")) r = ecb_bls12_381_g1_from_compressed_dev(ctx, di, in[0], max_n, 1, out[0], out[1], st);
            else if (!strcmp(op, "wei_mul")) r = ecb_wei_mul_dev(ctx, di, curve, in[0], in[1], max_n, out[0], out[1], st);
            else if (!strcmp(op, "wei_mul_base")) r = ecb_wei_mul_base_dev(ctx, di, curve, in[0], max_n, out[0], out[1], st);
            else if (!strcmp(op, "wei_decompress")) r = ecb_wei_decompress_dev(ctx, di, curve, in[0], in[1], max_n, out[0], out[1], st);
            else if (!strcmp(op, "ecdsa_verify_hashed")) r = ecb_ecdsa_verify_hashed_dev(ctx, di, curve, in[0], in[1], in[2], max_n, out[0], st);
            else if (!strcmp(op, "ecdsa_sign_hashed_vartime")) r = ecb_ecdsa_sign_hashed_vartime_dev(ctx, di, curve, in[0], in[1], in[2], max_n, out[0], out[1], st);
            else r = ecb_ecdsa_sign_hashed_dev(ctx, di, curve, in[0], in[1], in[2], max_n, out[0], out[1], st);
            g_warming = false;
            TRY(r);
            CU(cudaStreamSynchronize(d->stream));
            return ECB_OK;
        };
        rc = body();
        g_warming = false;
        cudaStreamSynchronize(d->stream);
        for (void* q : in) if (q) cudaFree(q);
        for (void* q : out) if (q) cudaFree(q);
        if (rc != ECB_OK) return rc;
    }
    return ECB_OK;
}
// what the context currently holds on device dev_index (0 when not built yet)
int ecb_get_info(ecb_ctx* ctx, int di, const char* key, long* value) {
    DevCtx* d = get_dev(ctx, di);
    if (!d || !key || !value) return ECB_ERR_INVALID_ARG;
    static const char* wn[4] = {"p256r1", "p384r1", "bls12_381_g1", "p256k1"};
    if (!strcmp(key, "ed25519_comb_w")) { *value = d->ed_w; return ECB_OK; }
    if (!strcmp(key, "ed25519_comb_windows")) { *value = d->ed_nwin; return ECB_OK; }
    if (!strcmp(key, "sm_count")) { *value = d->sm_count; return ECB_OK; }
    for (int c = 0; c < 4; c++) {
        if (!strcmp(key, (std::string(wn[c]) + "_comb_w").c_str())) { *value = d->wei_w[c]; return ECB_OK; }
        if (!strcmp(key, (std::string(wn[c]) + "_comb_windows").c_str())) { *value = d->wei_nwin[c]; return ECB_OK; }
    }
    return set_err(ctx, ECB_ERR_INVALID_ARG, std::string("ecb_get_info: unknown key ") + key);
}

}  // extern "C"
