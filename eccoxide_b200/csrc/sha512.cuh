// sha512.cuh — SHA-512 (FIPS 180-4) for the Ed25519 challenge k = SHA-512(R || A || M) mod l.
//
// Replaces, for the batched verification path, the reference's use of cryptoxide's SHA-512
// (src/protocol/ed25519.rs:8, :11-23 hash + reduce_wide_le, :139 in verify).  One thread hashes one
// message: the state is eight 64-bit words, the schedule a rolling window of 16; message bytes are
// fetched by a callback so the caller can concatenate R, A and a ragged message without copying.
#pragma once
#include "params_gen.cuh"

namespace ecb {

typedef unsigned long long w64;  // (limb.cuh's w64 is uint64_t: unsigned long on LP64)

ECB_DEV w64 rotr64(w64 x, int n) { return (x >> n) | (x << (64 - n)); }

struct Sha512 {
    w64 h[8];
    ECB_DEV void init(bool sha384 = false) {
        ECB_UNROLL
        for (int i = 0; i < 8; i++) h[i] = sha384 ? SHA384_H0[i] : SHA512_H0[i];
    }
    // one 128-byte block given as 16 big-endian words
    ECB_DEV void compress(w64* w) {
        w64 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        // 5 passes of 16 rounds, the 16 rounds unrolled: every index into the schedule window is a
        // compile-time constant, so w[] stays in registers (a rolled loop indexes it dynamically and
        // sends it to local memory: 5 local accesses per round)
        ECB_NOUNROLL
        for (int t0 = 0; t0 < 80; t0 += 16) {
            ECB_UNROLL
            for (int i = 0; i < 16; i++) {
                if (t0 != 0) {
                    w64 w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                    w64 s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7);
                    w64 s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
                    w[i] = w[i] + s0 + w[(i + 9) & 15] + s1;
                }
                w64 S1 = rotr64(e, 14) ^ rotr64(e, 18) ^ rotr64(e, 41);
                w64 ch = (e & f) ^ (~e & g);
                w64 t1 = hh + S1 + ch + SHA512_K[t0 + i] + w[i];
                w64 S0 = rotr64(a, 28) ^ rotr64(a, 34) ^ rotr64(a, 39);
                w64 mj = (a & b) ^ (a & c) ^ (b & c);
                w64 t2 = S0 + mj;
                hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
            }
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
};

// digest[0..64) = SHA-512 of the `len` bytes byte_at(0) .. byte_at(len-1); with sha384 the SHA-384
// initial value is used and the caller keeps the first 48 bytes
template <class BYTE_AT>
ECB_DEV void sha512_bytes(unsigned char* digest, size_t len, BYTE_AT byte_at, bool sha384 = false) {
    Sha512 st;
    st.init(sha384);
    w64 w[16];
    size_t nblocks = (len + 17 + 127) / 128;  // 0x80 marker + 16-byte length field
    ECB_NOUNROLL
    for (size_t blk = 0; blk < nblocks; blk++) {
        ECB_UNROLL
        for (int i = 0; i < 16; i++) {
            w64 v = 0;
            ECB_NOUNROLL
            for (int j = 0; j < 8; j++) {
                size_t pos = blk * 128 + (size_t)i * 8 + j;
                unsigned b = pos < len ? (unsigned)byte_at(pos) : (pos == len ? 0x80u : 0u);
                v = (v << 8) | b;
            }
            w[i] = v;
        }
        if (blk == nblocks - 1) {
            w[14] = (w64)(len >> 61);        // bit length, 128-bit big-endian
            w[15] = (w64)len << 3;
        }
        st.compress(w);
    }
    ECB_UNROLL
    for (int i = 0; i < 8; i++) {
        ECB_UNROLL
        for (int j = 0; j < 8; j++) digest[8 * i + j] = (unsigned char)(st.h[i] >> (56 - 8 * j));
    }
}

// The same hash with the first 4 NPW bytes given as NPW little-endian 32-bit words held in registers
// (pre(i), i a compile-time constant after unrolling: R || A of a signature, a seed, a nonce prefix)
// followed by the message bytes M[0, mlen): block 0 is assembled from registers, message words come
// from 8-byte loads when the address allows it.  Byte-for-byte the digest of sha512_bytes.
ECB_DEV u32 sha_bswap32(u32 x) { return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24); }
ECB_DEV w64 sha_msg_w64(const unsigned char* M, size_t mlen, size_t off) {
    if (off + 8 <= mlen) {
        const unsigned char* p = M + off;
        if ((reinterpret_cast<size_t>(p) & 7u) == 0) {
            w64 v = *reinterpret_cast<const w64*>(p);
            return ((w64)sha_bswap32((u32)v) << 32) | (w64)sha_bswap32((u32)(v >> 32));
        }
    }
    w64 v = 0;
    ECB_NOUNROLL
    for (int j = 0; j < 8; j++) {
        size_t pos = off + (size_t)j;
        unsigned b = pos < mlen ? (unsigned)M[pos] : (pos == mlen ? 0x80u : 0u);
        v = (v << 8) | b;
    }
    return v;
}
template <int NPW, class PRE>
ECB_DEV void sha512_prefixed(unsigned char* digest, PRE pre, const unsigned char* M, size_t mlen) {
    static_assert(NPW % 2 == 0 && NPW <= 28, "prefix must be whole 64-bit words inside the first block");
    const size_t len = 4 * (size_t)NPW + mlen;
    Sha512 st;
    st.init(false);
    w64 w[16];
    const size_t nblocks = (len + 17 + 127) / 128;
    ECB_UNROLL
    for (int i = 0; i < 16; i++) {
        if (i < NPW / 2) w[i] = ((w64)sha_bswap32(pre(2 * i)) << 32) | (w64)sha_bswap32(pre(2 * i + 1));
        else w[i] = sha_msg_w64(M, mlen, (size_t)(8 * i - 4 * NPW));
    }
    if (nblocks == 1) {
        w[14] = (w64)(len >> 61);
        w[15] = (w64)len << 3;
    }
    st.compress(w);
    ECB_NOUNROLL
    for (size_t blk = 1; blk < nblocks; blk++) {
        ECB_UNROLL
        for (int i = 0; i < 16; i++) w[i] = sha_msg_w64(M, mlen, blk * 128 + (size_t)(8 * i) - 4 * (size_t)NPW);
        if (blk == nblocks - 1) {
            w[14] = (w64)(len >> 61);
            w[15] = (w64)len << 3;
        }
        st.compress(w);
    }
    ECB_UNROLL
    for (int i = 0; i < 8; i++) {
        ECB_UNROLL
        for (int j = 0; j < 8; j++) digest[8 * i + j] = (unsigned char)(st.h[i] >> (56 - 8 * j));
    }
}

// ---- SHA-256 (ECDSA over p256r1: src/protocol/ecdsa.rs:288-292 hash_to_scalar) --------------------
ECB_DEV u32 rotr32(u32 x, int n) { return (x >> n) | (x << (32 - n)); }
template <class BYTE_AT>
ECB_DEV void sha256_bytes(unsigned char* digest, size_t len, BYTE_AT byte_at) {
    u32 h[8], w[16];
    ECB_UNROLL
    for (int i = 0; i < 8; i++) h[i] = SHA256_H0[i];
    size_t nblocks = (len + 9 + 63) / 64;
    ECB_NOUNROLL
    for (size_t blk = 0; blk < nblocks; blk++) {
        ECB_UNROLL
        for (int i = 0; i < 16; i++) {
            u32 v = 0;
            ECB_NOUNROLL
            for (int j = 0; j < 4; j++) {
                size_t pos = blk * 64 + (size_t)i * 4 + j;
                unsigned b = pos < len ? (unsigned)byte_at(pos) : (pos == len ? 0x80u : 0u);
                v = (v << 8) | b;
            }
            w[i] = v;
        }
        if (blk == nblocks - 1) {
            w[14] = (u32)((unsigned long long)len >> 29);
            w[15] = (u32)(len << 3);
        }
        u32 a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        ECB_NOUNROLL
        for (int t0 = 0; t0 < 64; t0 += 16) {      // as in Sha512::compress: static schedule indices
            ECB_UNROLL
            for (int i = 0; i < 16; i++) {
                if (t0 != 0) {
                    u32 w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
                    u32 s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
                    u32 s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
                    w[i] = w[i] + s0 + w[(i + 9) & 15] + s1;
                }
                u32 S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
                u32 ch = (e & f) ^ (~e & g);
                u32 t1 = hh + S1 + ch + SHA256_K[t0 + i] + w[i];
                u32 S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
                u32 mj = (a & b) ^ (a & c) ^ (b & c);
                u32 t2 = S0 + mj;
                hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
            }
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    ECB_UNROLL
    for (int i = 0; i < 8; i++) {
        digest[4 * i] = (unsigned char)(h[i] >> 24);
        digest[4 * i + 1] = (unsigned char)(h[i] >> 16);
        digest[4 * i + 2] = (unsigned char)(h[i] >> 8);
        digest[4 * i + 3] = (unsigned char)h[i];
    }
}

}  // namespace ecb
