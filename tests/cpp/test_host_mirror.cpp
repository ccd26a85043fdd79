// C++ host side over the C ABI (include/eccbatch.hpp), exercised the way the reference's own tests
// exercise its per-element API: the vectors come from tests/golden/reference_vectors.json (lifted
// from the reference's test modules by tools/extract_reference_vectors.py) and are handed over by
// tests/test_cpp_host.py as lines of "<name> <hex> <hex> ...".
//
//   usage: test_host_mirror <vector-file>      run every check on cuda:0, print "ok <n> checks"
//          test_host_mirror --no-device-check  expect the constructor to throw (no CPU fallback)
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "eccbatch.hpp"

using namespace eccoxide;

static int g_checks = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        g_checks++;                                                                  \
        if (!(cond)) {                                                               \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            std::exit(1);                                                            \
        }                                                                            \
    } while (0)

static std::vector<uint8_t> unhex(const std::string& h) {
    std::vector<uint8_t> out(h.size() / 2);
    for (size_t i = 0; i < out.size(); i++) out[i] = (uint8_t)std::stoi(h.substr(2 * i, 2), nullptr, 16);
    return out;
}
template <size_t N>
static Bytes<N> arr(const std::string& h) {
    std::vector<uint8_t> v = unhex(h);
    Bytes<N> a{};
    if (v.size() > N) { std::fprintf(stderr, "vector too long\n"); std::exit(2); }
    std::copy(v.begin(), v.end(), a.begin() + (N - v.size()));  // left-pad (big-endian KAT fields may be short)
    return a;
}
typedef std::vector<std::vector<std::string>> Rows;

int main(int argc, char** argv) {
    if (argc == 2 && std::string(argv[1]) == "--no-device-check") {
        try {
            Batch b;
        } catch (const BatchError& e) {
            std::printf("no device: BatchError code %d (%s)\n", e.code, e.what());
            return e.code == ECB_ERR_CUDA ? 0 : 1;
        }
        std::printf("a CUDA device is present\n");
        return 0;
    }
    if (argc != 2) return 2;
    std::map<std::string, Rows> V;
    std::ifstream f(argv[1]);
    for (std::string line; std::getline(f, line);) {
        std::istringstream ss(line);
        std::string name, tok;
        ss >> name;
        std::vector<std::string> row;
        while (ss >> tok) row.push_back(tok == "-" ? "" : tok);
        V[name].push_back(row);
    }
    Batch b;

    // x25519 (src/protocol/x25519.rs:118-160: RFC 7748 5.2 vectors and the 6.1 Diffie-Hellman exchange)
    {
        std::vector<Bytes<32>> k, u, want;
        for (auto& r : V["x25519"]) { k.push_back(arr<32>(r[0])); u.push_back(arr<32>(r[1])); want.push_back(arr<32>(r[2])); }
        CHECK(!k.empty() && x25519::x25519_batch(b, k, u) == want);
        auto& dh = V["x25519_dh"][0];
        auto pubs = x25519::x25519_base_batch(b, {arr<32>(dh[0]), arr<32>(dh[1])});
        auto shared = x25519::x25519_batch(b, {arr<32>(dh[0]), arr<32>(dh[1])}, {pubs[1], pubs[0]});
        CHECK(shared[0] == shared[1] && shared[0] == arr<32>(dh[2]));
    }
    // x448 (src/protocol/x448.rs:116-160)
    {
        std::vector<Bytes<56>> k, u, want;
        for (auto& r : V["x448"]) { k.push_back(arr<56>(r[0])); u.push_back(arr<56>(r[1])); want.push_back(arr<56>(r[2])); }
        CHECK(!k.empty() && x448::x448_batch(b, k, u) == want);
    }
    // ed25519 (src/protocol/ed25519.rs:271-290: RFC 8032 TEST 1-3): seed -> public key -> signature -> verify
    {
        std::vector<ed25519::SecretKey> seeds;
        std::vector<ed25519::PublicKey> pubs;
        std::vector<ed25519::Signature> sigs;
        std::vector<std::vector<uint8_t>> msgs;
        for (auto& r : V["ed25519"]) { seeds.push_back(arr<32>(r[0])); pubs.push_back(arr<32>(r[1])); msgs.push_back(unhex(r[2])); sigs.push_back(arr<64>(r[3])); }
        CHECK(!seeds.empty() && ed25519::public_key_batch(b, seeds) == pubs);
        CHECK(ed25519::sign_batch(b, seeds, msgs) == sigs);
        auto ok = ed25519::verify_batch(b, pubs, msgs, sigs);
        for (bool v : ok) CHECK(v);
        auto tampered = msgs;
        for (auto& m : tampered) m.push_back(0x21);
        for (bool v : ed25519::verify_batch(b, pubs, tampered, sigs)) CHECK(!v);
    }
    // edwards25519 mul_base / mul (src/curve/curve25519.rs:1374-1387 cross-algorithm identities):
    // k * B from the comb equals the variable-base product on the generator; a scalar >= l is refused
    {
        std::vector<curve25519::Scalar> k;
        for (auto& r : V["ed_scalar"]) k.push_back(arr<32>(r[0]));
        auto pts = curve25519::mul_base_batch(b, k);
        curve25519::Scalar one{};
        one[0] = 1;
        auto B = curve25519::mul_base_batch(b, {one})[0];
        CHECK(curve25519::mul_batch(b, std::vector<curve25519::PointAffine>(k.size(), B), k) == pts);
        curve25519::Scalar bad;
        bad.fill(0xff);
        k[1] = bad;
        try {
            curve25519::mul_base_batch(b, k);
            CHECK(false);
        } catch (const BatchError& e) {
            CHECK(e.code == ECB_ERR_NONCANONICAL_SCALAR && e.bad_index == 1);
        }
    }
    // NIST point-multiplication KATs (src/tests/kats.rs:35-38 over kats_data.rs): k * G, both algorithms
    {
        typedef weierstrass<P256R1> W;
        std::vector<W::Scalar> k;
        std::vector<W::PointAffine> want;
        for (auto& r : V["nist_p256"]) { k.push_back(arr<32>(r[0])); want.push_back(arr<64>(r[1] + r[2])); }
        auto got = W::mul_base_batch(b, k);
        CHECK(got.size() == want.size());
        for (size_t i = 0; i < got.size(); i++) CHECK(got[i].has_value() && *got[i] == want[i]);
        auto G = *W::mul_base_batch(b, {arr<32>("01")})[0];
        auto got2 = W::mul_batch(b, std::vector<W::PointAffine>(k.size(), G), k);
        for (size_t i = 0; i < got2.size(); i++) CHECK(got2[i].has_value() && *got2[i] == want[i]);
        // k = 0 -> the point at infinity (to_affine() == None); decompress recovers y from x and its parity
        CHECK(!W::mul_base_batch(b, {W::Scalar{}})[0].has_value());
        W::FieldElement x;
        std::copy(want[3].begin(), want[3].begin() + 32, x.begin());
        auto dec = W::decompress_batch(b, {x, x}, {(uint8_t)(want[3][63] & 1), (uint8_t)((want[3][63] & 1) ^ 1)});
        CHECK(dec[0].has_value() && *dec[0] == want[3] && dec[1].has_value() && *dec[1] != want[3]);
        // an off-curve point is refused with its index (PointAffine::from_coordinate -> None)
        auto offc = G;
        offc[63] ^= 1;
        try {
            W::mul_batch(b, {G, offc}, {k[0], k[1]});
            CHECK(false);
        } catch (const BatchError& e) {
            CHECK(e.code == ECB_ERR_POINT_NOT_ON_CURVE && e.bad_index == 1);
        }
    }
    {
        typedef weierstrass<P384R1> W;
        std::vector<W::Scalar> k;
        std::vector<W::PointAffine> want;
        for (auto& r : V["nist_p384"]) { k.push_back(arr<48>(r[0])); want.push_back(arr<96>(r[1] + r[2])); }
        auto got = W::mul_base_batch(b, k);
        for (size_t i = 0; i < got.size(); i++) CHECK(got[i].has_value() && *got[i] == want[i]);
    }
    // ECDSA (src/protocol/ecdsa.rs:808-878: RFC 6979 A.2.5 with the listed nonces): sign, verify, tamper
    {
        typedef ecdsa<P256R1> E;
        typedef weierstrass<P256R1> W;
        std::vector<W::Scalar> d, k, z;
        std::vector<E::Signature> want;
        for (auto& r : V["ecdsa_p256"]) { d.push_back(arr<32>(r[0])); k.push_back(arr<32>(r[1])); z.push_back(arr<32>(r[2])); want.push_back(arr<64>(r[3] + r[4])); }
        auto sig = E::sign_hashed_batch(b, d, k, z);
        std::vector<E::Signature> sigs;
        for (size_t i = 0; i < sig.size(); i++) { CHECK(sig[i].has_value() && *sig[i] == want[i]); sigs.push_back(*sig[i]); }
        auto Q = W::mul_base_batch(b, d);
        std::vector<W::PointAffine> q;
        for (auto& p : Q) q.push_back(*p);
        for (bool v : E::verify_hashed_batch(b, q, z, sigs)) CHECK(v);
        auto z2 = z;
        for (auto& s : z2) s[31] ^= 1;
        for (bool v : E::verify_hashed_batch(b, q, z2, sigs)) CHECK(!v);
        CHECK(!E::sign_hashed_batch(b, {W::Scalar{}}, {k[0]}, {z[0]})[0].has_value());  // zero secret: no signature
    }
    // BLS12-381 G1 (src/curve/bls12_381/g1.rs:605-680 serialization KATs, :313-368 OFF_SUBGROUP)
    {
        typedef weierstrass<Bls12381G1> W;
        std::vector<W::Scalar> k;
        std::vector<bls12_381::g1::Compressed> enc;
        for (auto& r : V["bls_compressed"]) { k.push_back(arr<32>(r[0])); enc.push_back(arr<48>(r[1])); }
        auto pts = W::mul_base_batch(b, k);
        CHECK(bls12_381::g1::to_compressed_batch(b, pts) == enc);
        auto back = bls12_381::g1::from_compressed_batch(b, enc, true);
        for (size_t i = 0; i < back.size(); i++) CHECK(back[i].has_value() && *back[i] == *pts[i]);
        std::vector<bls12_381::g1::Compressed> off;
        std::vector<bls12_381::g1::PointAffine> unc;
        for (auto& r : V["bls_off_subgroup"]) { off.push_back(arr<48>(r[0])); unc.push_back(arr<96>(r[1])); }
        auto loose = bls12_381::g1::from_compressed_batch(b, off, false);
        auto strict = bls12_381::g1::from_compressed_batch(b, off, true);
        for (size_t i = 0; i < off.size(); i++) CHECK(loose[i].has_value() && *loose[i] == unc[i] && !strict[i].has_value());
        CHECK(bls12_381::g1::to_compressed_batch(b, {std::nullopt})[0][0] == 0xc0);
    }
    std::printf("ok %d checks, %llu kernel launches\n", g_checks, ecb_launch_count(b.handle()));
    return 0;
}
