// fused.cuh — small-batch form of the fixed-base comb: lanes cooperate on one scalar and the affine
// conversion happens inside the same launch (device code only: warp shuffles + shared memory).
//
// Why: BASELINE configs[0] is 2^16 scalars = 443 per SM.  With one thread per scalar and a second
// kernel for the batch inversion (the large-batch form, kernels.cuh) that batch is latency-bound twice:
// 14 warps per SM walk ten dependent window additions, the projective planes go out to HBM, and a
// second launch pays the latency of its two passes around one inversion.  Here
//   * the nwin window additions of a scalar — independent in the reference too (curve25519.rs:844-849:
//     `acc = acc + table[i][digit_i]` for 64 windows) — are split over LANES adjacent lanes, each lane
//     accumulating the windows i = lane (mod LANES); the partial sums are added with a shuffle
//     butterfly (one complete extended addition per level, Point::add curve25519.rs:695);
//   * every thread then holds one denominator and the block inverts them TOGETHER: inclusive prefix and
//     suffix products over the lanes of each warp by shuffles, the warp totals through shared memory,
//     ONE safegcd inversion per block, two products per thread to peel its own inverse
//     (Montgomery's trick as a product scan; replaces to_affine's per-point Fermat inverse,
//     curve25519.rs:663, :155-200);
//   * the finisher multiplies and writes the reference's wire bytes straight from registers: the
//     projective planes never leave the SM.
// The host picks LANES and the block size so that the whole batch is ONE wave of at most 512 threads
// per SM (tu_ed25519.cu: fused_shape).
#pragma once
#include "kernels.cuh"

namespace ecb {

template <class FT>
__device__ __forceinline__ void fe_shfl_up(typename FT::el& r, const typename FT::el& a, unsigned d) {
#pragma unroll
    for (int k = 0; k < FT::N; k++) r.v[k] = __shfl_up_sync(0xffffffffu, a.v[k], d);
}
template <class FT>
__device__ __forceinline__ void fe_shfl_down(typename FT::el& r, const typename FT::el& a, unsigned d) {
#pragma unroll
    for (int k = 0; k < FT::N; k++) r.v[k] = __shfl_down_sync(0xffffffffu, a.v[k], d);
}
template <class FT>
__device__ __forceinline__ void fe_shfl_idx(typename FT::el& r, const typename FT::el& a, int src) {
#pragma unroll
    for (int k = 0; k < FT::N; k++) r.v[k] = __shfl_sync(0xffffffffu, a.v[k], src);
}
template <class FT>
__device__ __forceinline__ void fe_shfl_xor(typename FT::el& r, const typename FT::el& a, int m) {
#pragma unroll
    for (int k = 0; k < FT::N; k++) r.v[k] = __shfl_xor_sync(0xffffffffu, a.v[k], m);
}

// Block-cooperative inversion: every WORK thread of the block passes one element z and receives z^-1
// (z = 0 -> treated as 1, `zero` set).  The block is nww <= FUSED_MAXW - 1 work warps plus ONE more warp
// (the last) that holds no element and does nothing but the inversion, so that everything which does
// not need the inverse happens beside it instead of before or after it:
//   all work warps   inclusive prefix products over their lanes; lane 31 publishes the warp total   | barrier
//   inversion warp   product of the warp totals (butterfly), ONE inverse by the 32 lanes (sg_modinv_warp)
//   work warps       meanwhile: suffix products, the product of the OTHER warps' totals, and from those
//                    E = product of every other element of the block                                  | barrier
//   work threads     z^-1 = E * inverse: one product after the inverse is known.
// sh: (FUSED_MAXW + 1) * N words of shared memory; jump: the safegcd jump table (modinv.cuh) staged in shared memory.
#define FUSED_MAXW 16
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
template <class FT>
__device__ __forceinline__ void block_invert(typename FT::el& zinv, typename FT::el z, u32& zero, u32* sh, const u32* jump) {
    typedef typename FT::el fe;
    constexpr int N = FT::N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nww = (int)(blockDim.x >> 5) - 1;
    const bool inv_role = warp >= nww;
    u32* sh_tot = sh;
    u32* sh_inv = sh + FUSED_MAXW * N;
    fe one;
    FT::set_one(one);
    zero = FT::is_zero(z);
    FT::select(z, zero, one, z);
    fe P = z, o, m;
    if (!inv_role) {
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            fe_shfl_up<FT>(o, P, d);
            FT::mul(m, P, o);
            if (lane >= d) P = m;
        }
        if (lane == 31) {
#pragma unroll
            for (int k = 0; k < N; k++) sh_tot[k * FUSED_MAXW + warp] = P.v[k];
        }
    }
    __syncthreads();
    // both halves of a warp load the same 16 slots, so four butterfly levels leave the product in every lane
    fe t = one, E = one;
    const int slot = lane & 15;
    if (slot < nww && (inv_role || slot != warp)) {
#pragma unroll
        for (int k = 0; k < N; k++) t.v[k] = sh_tot[k * FUSED_MAXW + slot];
    }
    if (inv_role) {
#pragma unroll 1
        for (int d = 1; d < 16; d <<= 1) {
            fe_shfl_xor<FT>(o, t, d);
            FT::mul(t, t, o);
        }
        FT::invert_warp(m, t, jump);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < N; k++) sh_inv[k] = m.v[k];
        }
    } else {
        fe S = z, exP, exS;
#pragma unroll 1
        for (int d = 1; d < 32; d <<= 1) {
            fe_shfl_down<FT>(o, S, d);
            FT::mul(m, S, o);
            if (lane + d < 32) S = m;
        }
#pragma unroll 1
        for (int d = 1; d < 16; d <<= 1) {   // product of the other warps' totals
            fe_shfl_xor<FT>(o, t, d);
            FT::mul(t, t, o);
        }
        fe_shfl_up<FT>(exP, P, 1);
        fe_shfl_down<FT>(exS, S, 1);
        if (lane == 0) exP = one;
        if (lane == 31) exS = one;
        FT::mul(m, exP, exS);
        FT::mul(E, m, t);
    }
    __syncthreads();
    fe tinv;
#pragma unroll
    for (int k = 0; k < N; k++) tinv.v[k] = sh_inv[k];
    FT::mul(zinv, E, tinv);
    __syncthreads();   // sh is reused by the next tile
}

// finishers of the fused kernels: den() names the element to invert, operator() writes the wire bytes
struct FusedEdXY {          // Point::to_affine + to_bytes (curve25519.rs:663): x_le || y_le
    u32* out;
    __device__ __forceinline__ void den(fe25519& d, const ge_p3& p) const { d = p.Z; }
    __device__ __forceinline__ void operator()(size_t idx, const ge_p3& p, const fe25519& dinv, u32) const {
        fe25519 x, y;
        F::mul(x, p.X, dinv);
        F::mul(y, p.Y, dinv);
        F::freeze(x, x);
        F::freeze(y, y);
        st_words<8>(out + idx * 16, x.v);
        st_words<8>(out + idx * 16 + 8, y.v);
    }
};
struct FusedEdCompressed {  // encode_point (protocol/ed25519.rs:27); element idx at out + idx * stride words
    u32* out; size_t stride;
    __device__ __forceinline__ void den(fe25519& d, const ge_p3& p) const { d = p.Z; }
    __device__ __forceinline__ void operator()(size_t idx, const ge_p3& p, const fe25519& dinv, u32) const {
        fe25519 x, y;
        F::mul(x, p.X, dinv);
        F::mul(y, p.Y, dinv);
        F::freeze(x, x);
        F::freeze(y, y);
        y.v[7] |= (x.v[0] & 1u) << 31;
        st_words<8>(out + idx * stride, y.v);
    }
};
struct FusedEdMontU {       // x25519_base: u = (Z + Y) / (Z - Y), 0 when Z = Y (x25519.rs:49, curve25519.rs:512)
    u32* out;
    __device__ __forceinline__ void den(fe25519& d, const ge_p3& p) const { F::sub(d, p.Z, p.Y); }
    __device__ __forceinline__ void operator()(size_t idx, const ge_p3& p, const fe25519& dinv, u32 zero) const {
        fe25519 s, u;
        F::add(s, p.Z, p.Y);
        F::mul(u, s, dinv);
        F::freeze(u, u);
        u32 m = zero ? 0u : 0xffffffffu;
#pragma unroll
        for (int i = 0; i < 8; i++) u.v[i] &= m;
        st_words<8>(out + idx * 8, u.v);
    }
};

// One tile = (blockDim.x - 32) / LANES scalars (the last warp is the inversion warp of block_invert); a block
// walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ...
template <int LANES, bool CLAMP, class FIN>
__device__ __forceinline__ void ed25519_mul_base_fused_block(size_t n, const u32* scalars, const u32* table, int W, int nwin,
                                                             int stride, FIN fin, unsigned long long* status, u32* sh,
                                                             u32* jump, unsigned long long* trace) {
    const int tid = threadIdx.x;
    sg_stage_jump_table(jump);   // read after the first __syncthreads() of block_invert, long after these loads
    if (trace && tid == 0) trace[(size_t)blockIdx.x * 4 + 0] = globaltimer_ns();
    const int work = (int)blockDim.x - 32;
    const int l = tid % LANES;
    const size_t per_tile = work / LANES;
    for (size_t tile = blockIdx.x; tile * per_tile < n; tile += gridDim.x) {
        const size_t idx = tile * per_tile + tid / LANES;
        const bool live = tid < work && idx < n;
        ge_p3 acc;
        if (live) {
            u32 k[9];
            ed25519_load_scalar<CLAMP>(k, idx, scalars, l == 0 ? status : nullptr);
            ed25519_comb_partial(acc, k, table, W, nwin, stride, l, LANES, LANES > 1);
        } else {
            ge_identity(acc);
        }
        if (tid < work) {
#pragma unroll 1
            for (int d = 1; d < LANES; d <<= 1) {   // butterfly: afterwards every lane of the group holds the sum
                ge_p3 o;
                fe_shfl_xor<F25519>(o.X, acc.X, d);
                fe_shfl_xor<F25519>(o.Y, acc.Y, d);
                fe_shfl_xor<F25519>(o.Z, acc.Z, d);
                fe_shfl_xor<F25519>(o.T, acc.T, d);
                ge_add_p3<true>(acc, acc, o);
            }
        }
        fe25519 den, dinv;
        F::set_one(den);
        if (live && l == 0) fin.den(den, acc);
        u32 zero;
        if (trace && tid == 0) trace[(size_t)blockIdx.x * 4 + 1] = globaltimer_ns();
        block_invert<F25519>(dinv, den, zero, sh, jump);
        if (trace && tid == 0) trace[(size_t)blockIdx.x * 4 + 2] = globaltimer_ns();
        if (live && l == 0) fin(idx, acc, dinv, zero);
    }
    if (trace && tid == 0) trace[(size_t)blockIdx.x * 4 + 3] = globaltimer_ns();
}

}  // namespace ecb
