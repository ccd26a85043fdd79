// eccbatch.hpp — C++17 host-side mirror of the reference's per-curve API over the C ABI (eccbatch.h).
//
// The reference is a Rust crate; its toolchain is not in this image, so the compiled host side that
// runs here is this header (the Rust siblings a maintainer would add are in bindings/rust/).  Names
// follow the reference's modules and functions; every batch function is element-wise the function
// it names, on the reference's own wire encodings (to_bytes / from_bytes):
//
//   eccoxide::curve25519::mul_base_batch     Point::mul_base            src/curve/curve25519.rs:840
//   eccoxide::curve25519::mul_batch          &Point * &Scalar           src/curve/curve25519.rs:1274
//   eccoxide::x25519::x25519_batch           x25519::x25519             src/protocol/x25519.rs:36
//   eccoxide::x25519::x25519_base_batch      x25519::x25519_base        src/protocol/x25519.rs:49
//   eccoxide::x448::x448_batch               x448::x448                 src/protocol/x448.rs:34
//   eccoxide::weierstrass<C>::mul_batch      &Point * &Scalar           src/curve/fiat/curve_macros.rs:321
//   eccoxide::weierstrass<C>::mul_base_batch Point::mul_base            src/curve/fiat/curve_macros.rs:55
//   eccoxide::weierstrass<C>::decompress_batch PointAffine::decompress  src/curve/affine.rs:48
//   eccoxide::ecdsa<C>::verify_hashed_batch  ecdsa::verify_hashed       src/protocol/ecdsa.rs:205
//   eccoxide::ecdsa<C>::sign_hashed_batch    ecdsa::sign_hashed         src/protocol/ecdsa.rs:165
//   eccoxide::ed25519::verify_batch          PublicKey::verify          src/protocol/ed25519.rs:119
//   eccoxide::ed25519::public_key_batch      SecretKey::public_key      src/protocol/ed25519.rs:81
//   eccoxide::ed25519::sign_batch            SecretKey::sign            src/protocol/ed25519.rs:112
//   eccoxide::bls12_381::g1::from_compressed_batch / to_compressed_batch  src/curve/bls12_381/serialize.rs:286, :400
//   eccoxide::bls12_381::g1::mul_batch_in_subgroup   &Point * &Scalar on G1 src/curve/bls12_381/g1.rs:105 (endomorphism)
//
// Error behaviour: where the reference's per-element constructors return None (Scalar::from_bytes on
// a value >= the order, PointAffine::from_coordinate off the curve) the batch call throws BatchError
// carrying the index of the first offending element; results that are Option in the reference
// (to_affine() of the identity, decompress, from_compressed, sign_hashed) come back as std::optional.
// There is no CPU fallback: without a CUDA device the Batch constructor throws.
// (Calls that report an index run before Batch::check reads it: argument evaluation order is unspecified.)
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "eccbatch.h"

namespace eccoxide {

struct BatchError : std::runtime_error {
    int code;
    size_t bad_index;  // (size_t)-1 when the failure is not tied to an element
    BatchError(int c, const std::string& msg, size_t bad) : std::runtime_error(msg), code(c), bad_index(bad) {}
};

template <size_t N>
using Bytes = std::array<uint8_t, N>;

// One context over the given CUDA devices (empty: device 0).  Batches are sharded by contiguous slice.
class Batch {
  public:
    explicit Batch(const std::vector<int>& devices = {}) {
        int rc = ecb_init(devices.empty() ? nullptr : devices.data(), (int)devices.size(), &ctx_);
        if (rc != ECB_OK) throw BatchError(rc, "ecb_init failed: no CUDA device or out of memory (there is no CPU fallback)", (size_t)-1);
    }
    ~Batch() { ecb_destroy(ctx_); }
    Batch(const Batch&) = delete;
    Batch& operator=(const Batch&) = delete;
    ecb_ctx* handle() const { return ctx_; }
    void set_option(const char* key, long value) const { check(ecb_set_option(ctx_, key, value)); }
    void check(int rc, size_t bad = (size_t)-1) const {
        if (rc != ECB_OK) throw BatchError(rc, ecb_last_error(ctx_), bad);
    }

  private:
    ecb_ctx* ctx_ = nullptr;
};

namespace detail {
template <size_t N>
const uint8_t* flat(const std::vector<Bytes<N>>& v) { return v.empty() ? nullptr : v[0].data(); }
template <size_t N>
uint8_t* flat(std::vector<Bytes<N>>& v) { return v.empty() ? nullptr : v[0].data(); }
inline void same(size_t a, size_t b) {
    if (a != b) throw BatchError(ECB_ERR_INVALID_ARG, "length mismatch", (size_t)-1);
}
// ragged messages -> one blob + n + 1 offsets
struct Ragged {
    std::vector<uint8_t> blob;
    std::vector<uint64_t> off;
    explicit Ragged(const std::vector<std::vector<uint8_t>>& msgs) {
        off.reserve(msgs.size() + 1);
        off.push_back(0);
        for (const auto& m : msgs) {
            blob.insert(blob.end(), m.begin(), m.end());
            off.push_back(blob.size());
        }
        blob.push_back(0);
    }
};
}  // namespace detail

// ---- edwards25519 (src/curve/curve25519.rs) -----------------------------------------------------
namespace curve25519 {
using Scalar = Bytes<32>;       // Scalar::to_bytes_le, canonical (< l)
using PointAffine = Bytes<64>;  // to_affine(): x || y, little-endian canonical
inline std::vector<PointAffine> mul_base_batch(const Batch& b, const std::vector<Scalar>& k) {
    std::vector<PointAffine> out(k.size());
    size_t bad = (size_t)-1;
    int rc = ecb_ed25519_mul_base(b.handle(), detail::flat(k), k.size(), detail::flat(out), &bad);
    b.check(rc, bad);
    return out;
}
inline std::vector<PointAffine> mul_batch(const Batch& b, const std::vector<PointAffine>& p, const std::vector<Scalar>& k) {
    detail::same(p.size(), k.size());
    std::vector<PointAffine> out(k.size());
    size_t bad = (size_t)-1;
    int rc = ecb_ed25519_mul(b.handle(), detail::flat(k), detail::flat(p), k.size(), detail::flat(out), &bad);
    b.check(rc, bad);
    return out;
}
}  // namespace curve25519

namespace x25519 {
inline std::vector<Bytes<32>> x25519_batch(const Batch& b, const std::vector<Bytes<32>>& k, const std::vector<Bytes<32>>& u) {
    detail::same(k.size(), u.size());
    std::vector<Bytes<32>> out(k.size());
    b.check(ecb_x25519(b.handle(), detail::flat(k), detail::flat(u), k.size(), detail::flat(out)));
    return out;
}
inline std::vector<Bytes<32>> x25519_base_batch(const Batch& b, const std::vector<Bytes<32>>& k) {
    std::vector<Bytes<32>> out(k.size());
    b.check(ecb_x25519_base(b.handle(), detail::flat(k), k.size(), detail::flat(out)));
    return out;
}
}  // namespace x25519

namespace x448 {
inline std::vector<Bytes<56>> x448_batch(const Batch& b, const std::vector<Bytes<56>>& k, const std::vector<Bytes<56>>& u) {
    detail::same(k.size(), u.size());
    std::vector<Bytes<56>> out(k.size());
    b.check(ecb_x448(b.handle(), detail::flat(k), detail::flat(u), k.size(), detail::flat(out)));
    return out;
}
}  // namespace x448

// ---- short Weierstrass curves (src/curve/sec2/p256r1.rs, p384r1.rs, src/curve/bls12_381/g1.rs) ------
struct P256R1 { static constexpr int id = ECB_CURVE_P256R1; static constexpr size_t FB = 32, SB = 32; };
struct P384R1 { static constexpr int id = ECB_CURVE_P384R1; static constexpr size_t FB = 48, SB = 48; };
struct Bls12381G1 { static constexpr int id = ECB_CURVE_BLS12_381_G1; static constexpr size_t FB = 48, SB = 32; };

template <class C>
struct weierstrass {
    using Scalar = Bytes<C::SB>;            // big-endian, canonical (< n)
    using FieldElement = Bytes<C::FB>;      // big-endian, canonical (< p)
    using PointAffine = Bytes<2 * C::FB>;   // x || y
    using MaybePoint = std::optional<PointAffine>;  // nullopt = the point at infinity (to_affine() == None)

    static std::vector<MaybePoint> wrap(std::vector<PointAffine>& xy, const std::vector<uint8_t>& none) {
        std::vector<MaybePoint> out(xy.size());
        for (size_t i = 0; i < xy.size(); i++)
            if (!none[i]) out[i] = xy[i];
        return out;
    }
    static std::vector<MaybePoint> mul_batch(const Batch& b, const std::vector<PointAffine>& p, const std::vector<Scalar>& k) {
        detail::same(p.size(), k.size());
        std::vector<PointAffine> xy(k.size());
        std::vector<uint8_t> inf(k.size());
        size_t bad = (size_t)-1;
        int rc = ecb_wei_mul(b.handle(), C::id, detail::flat(k), detail::flat(p), nullptr, k.size(), detail::flat(xy), inf.data(), &bad);
    b.check(rc, bad);
        return wrap(xy, inf);
    }
    static std::vector<MaybePoint> mul_base_batch(const Batch& b, const std::vector<Scalar>& k) {
        std::vector<PointAffine> xy(k.size());
        std::vector<uint8_t> inf(k.size());
        size_t bad = (size_t)-1;
        int rc = ecb_wei_mul_base(b.handle(), C::id, detail::flat(k), k.size(), detail::flat(xy), inf.data(), &bad);
    b.check(rc, bad);
        return wrap(xy, inf);
    }
    // sign[i]: 0 = Sign::Positive (even y), 1 = Sign::Negative (odd y)
    static std::vector<MaybePoint> decompress_batch(const Batch& b, const std::vector<FieldElement>& x, const std::vector<uint8_t>& sign) {
        detail::same(x.size(), sign.size());
        std::vector<PointAffine> xy(x.size());
        std::vector<uint8_t> ok(x.size());
        b.check(ecb_wei_decompress(b.handle(), C::id, detail::flat(x), sign.data(), x.size(), detail::flat(xy), ok.data()));
        for (auto& o : ok) o = !o;
        return wrap(xy, ok);
    }
};

template <class C>
struct ecdsa {
    using W = weierstrass<C>;
    using Signature = Bytes<2 * C::SB>;  // r || s big-endian
    static std::vector<bool> verify_hashed_batch(const Batch& b, const std::vector<typename W::PointAffine>& q,
                                                 const std::vector<typename W::Scalar>& z, const std::vector<Signature>& sig) {
        detail::same(q.size(), z.size());
        detail::same(q.size(), sig.size());
        std::vector<uint8_t> ok(q.size());
        size_t bad = (size_t)-1;
        int rc = ecb_ecdsa_verify_hashed(b.handle(), C::id, detail::flat(q), detail::flat(z), detail::flat(sig), q.size(), ok.data(), &bad);
    b.check(rc, bad);
        return std::vector<bool>(ok.begin(), ok.end());
    }
    // constant-time kernels on the device; the ecb_*_vartime entry points are the fast forms for public data
    static std::vector<std::optional<Signature>> sign_hashed_batch(const Batch& b, const std::vector<typename W::Scalar>& secret,
                                                                   const std::vector<typename W::Scalar>& nonce,
                                                                   const std::vector<typename W::Scalar>& hashed) {
        detail::same(secret.size(), nonce.size());
        detail::same(secret.size(), hashed.size());
        std::vector<Signature> rs(secret.size());
        std::vector<uint8_t> ok(secret.size());
        b.check(ecb_ecdsa_sign_hashed(b.handle(), C::id, detail::flat(secret), detail::flat(nonce), detail::flat(hashed), secret.size(),
                                      detail::flat(rs), ok.data()));
        std::vector<std::optional<Signature>> out(rs.size());
        for (size_t i = 0; i < rs.size(); i++)
            if (ok[i]) out[i] = rs[i];
        return out;
    }
};

// ---- Ed25519 (src/protocol/ed25519.rs) --------------------------------------------------------------
namespace ed25519 {
using PublicKey = Bytes<32>;
using SecretKey = Bytes<32>;  // the seed
using Signature = Bytes<64>;
inline std::vector<bool> verify_batch(const Batch& b, const std::vector<PublicKey>& pk, const std::vector<std::vector<uint8_t>>& msgs,
                                      const std::vector<Signature>& sig) {
    detail::same(pk.size(), msgs.size());
    detail::same(pk.size(), sig.size());
    detail::Ragged r(msgs);
    std::vector<uint8_t> ok(pk.size());
    b.check(ecb_ed25519_verify(b.handle(), detail::flat(pk), r.blob.data(), r.off.data(), detail::flat(sig), pk.size(), ok.data()));
    return std::vector<bool>(ok.begin(), ok.end());
}
// constant-time kernels on the device; the ecb_*_vartime entry points are the fast forms for public data
inline std::vector<PublicKey> public_key_batch(const Batch& b, const std::vector<SecretKey>& seeds) {
    std::vector<PublicKey> out(seeds.size());
    b.check(ecb_ed25519_public_from_seed(b.handle(), detail::flat(seeds), seeds.size(), detail::flat(out)));
    return out;
}
inline std::vector<Signature> sign_batch(const Batch& b, const std::vector<SecretKey>& seeds, const std::vector<std::vector<uint8_t>>& msgs) {
    detail::same(seeds.size(), msgs.size());
    detail::Ragged r(msgs);
    std::vector<Signature> out(seeds.size());
    b.check(ecb_ed25519_sign(b.handle(), detail::flat(seeds), nullptr, r.blob.data(), r.off.data(), seeds.size(), detail::flat(out)));
    return out;
}
}  // namespace ed25519

// ---- BLS12-381 G1 standard encodings (src/curve/bls12_381/serialize.rs) -------------------------------
namespace bls12_381 {
namespace g1 {
using PointAffine = Bytes<96>;
using Compressed = Bytes<48>;
// from_compressed (check_subgroup) / from_compressed_oncurve_only (!check_subgroup)
inline std::vector<std::optional<PointAffine>> from_compressed_batch(const Batch& b, const std::vector<Compressed>& enc, bool check_subgroup = true) {
    std::vector<PointAffine> xy(enc.size());
    std::vector<uint8_t> ok(enc.size());
    b.check(ecb_bls12_381_g1_from_compressed(b.handle(), detail::flat(enc), enc.size(), check_subgroup ? 1 : 0, detail::flat(xy), ok.data()));
    std::vector<std::optional<PointAffine>> out(enc.size());
    for (size_t i = 0; i < enc.size(); i++)
        if (ok[i]) out[i] = xy[i];
    return out;
}
// Point::to_compressed: nullopt = the identity (0xc0 00 .. 00)
inline std::vector<Compressed> to_compressed_batch(const Batch& b, const std::vector<std::optional<PointAffine>>& p) {
    std::vector<PointAffine> xy(p.size());
    std::vector<uint8_t> inf(p.size());
    for (size_t i = 0; i < p.size(); i++) {
        inf[i] = !p[i].has_value();
        xy[i] = p[i].value_or(PointAffine{});
    }
    std::vector<Compressed> out(p.size());
    b.check(ecb_bls12_381_g1_to_compressed(b.handle(), detail::flat(xy), inf.data(), p.size(), detail::flat(out)));
    return out;
}
// &Point * &Scalar for points KNOWN to be in the prime-order subgroup (outputs of from_compressed_batch(.., true) or
// of mul_base_batch): the device splits the scalar over the endomorphism of the subgroup test (g1.rs:105) — the same
// group element on G1, 1.35x faster; for a point outside G1 the result would not be k * P, which is why the plain
// weierstrass<Bls12381G1>::mul_batch is the sibling of Point::mul.  The option is scoped to the call.
inline std::vector<std::optional<PointAffine>> mul_batch_in_subgroup(const Batch& b, const std::vector<PointAffine>& p,
                                                                     const std::vector<Bytes<32>>& k) {
    struct Scope {
        const Batch& b;
        explicit Scope(const Batch& bb) : b(bb) { b.set_option("bls12_381_g1_glv", 1); }
        ~Scope() { ecb_set_option(b.handle(), "bls12_381_g1_glv", 0); }
    } scope(b);
    return weierstrass<Bls12381G1>::mul_batch(b, p, k);
}
}  // namespace g1
}  // namespace bls12_381

}  // namespace eccoxide
