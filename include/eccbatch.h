/* eccbatch.h — C ABI of libeccbatch: batched elliptic-curve scalar multiplication on B200 (sm_100a).
 *
 * This is the drop-in boundary for the one data-parallel hot path of vincenthz/eccoxide: many
 * independent scalar multiplications.  The reference has no FFI of its own (pure Rust, one element
 * per call); each entry point below is the batch sibling of one public per-element function and
 * reproduces that function's result byte for byte on every element.  File:line citations are
 * relative to the reference tree.  The Rust shim that binds these is in bindings/rust/ and
 * INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 (ECB_OK) or a negative ECB_ERR_* code; ecb_last_error() gives text
 *   - all buffers are caller-owned, contiguous, array-of-structures, in the reference's own wire
 *     encodings (curve25519: little-endian; Weierstrass curves: big-endian), element i at i*size
 *   - host entry points copy host->device, run the kernels and copy back before returning
 *     (blocking); *_dev entry points take device pointers on device `dev_index` of the context and
 *     enqueue on `stream` (a cudaStream_t) without synchronising and WITHOUT allocating: the generator
 *     comb and the work buffers they need are provided beforehand by ecb_warm(ctx, op, curve, max_n);
 *     a *_dev call that finds them missing or too small fails with ECB_ERR_NOT_READY; n == 0 is a
 *     no-op; device pointers must be 16-byte aligned (the kernels move elements with 128-bit loads
 *     and stores; cudaMalloc gives 256); successive *_dev calls share one set of work buffers, so
 *     enqueue them on ONE stream or order the streams with events
 *   - a batch is sharded by contiguous slice over the devices of the context; no collective
 *   - threading: host entry points may be called concurrently on one context (calls are serialised
 *     per device); the *_dev entry points use the context's slot-0 work buffers without locking, so
 *     issue them from one thread per device and do not mix them with concurrent host calls
 *   - input validation mirrors the reference's Option/CtOption results: a non-canonical scalar
 *     (Scalar::from_bytes -> None, src/curve/fiat/field_macros.rs:645) or an off-curve /
 *     non-canonical point (PointAffine::from_coordinate -> None, src/curve/affine.rs:77;
 *     Point::from_coordinate, src/curve/curve25519.rs:649) fails the whole call with
 *     ECB_ERR_NONCANONICAL_SCALAR / ECB_ERR_POINT_NOT_ON_CURVE and *bad_index = first offender
 *   - there is no CPU fallback: without a CUDA device every call fails with ECB_ERR_CUDA
 *   - the scalar-multiplication and verification entry points are NOT constant time (indexed table loads,
 *     data-dependent skips): they are meant for public data.  Key generation and signing, and
 *     ecb_ed25519_mul_base_ct, run constant-time kernels (csrc/ct.cuh) unless their *_vartime form is
 *     called; see DESIGN.md section 8
 */
#ifndef ECCBATCH_H
#define ECCBATCH_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECB_OK 0
#define ECB_ERR_CUDA (-1)
#define ECB_ERR_INVALID_ARG (-2)
#define ECB_ERR_NONCANONICAL_SCALAR (-3)
#define ECB_ERR_POINT_NOT_ON_CURVE (-4)
#define ECB_ERR_OOM (-5)
#define ECB_ERR_NOT_READY (-6) /* a *_dev call found a table or work buffer missing: ecb_warm() first */

/* Weierstrass curve ids (src/curve/sec2/p256r1.rs, p384r1.rs, src/curve/bls12_381/g1.rs) */
#define ECB_CURVE_P256R1 0       /* field 32 B, scalar 32 B */
#define ECB_CURVE_P384R1 1       /* field 48 B, scalar 48 B */
#define ECB_CURVE_BLS12_381_G1 2 /* field 48 B, scalar 32 B */
#define ECB_CURVE_P256K1 3       /* secp256k1 (src/curve/sec2/p256k1.rs): field 32 B, scalar 32 B; mul, mul_base, decompress */

typedef struct ecb_ctx ecb_ctx;

/* Create a context over `n_dev` CUDA devices (device_ids == NULL, n_dev == 0: device 0 only): streams,
 * status words, nothing else.  The per-device generator comb tables — the role of the reference's
 * OnceLock'd generator_comb() (src/curve/curve25519.rs:881, src/curve/fiat/curve_macros.rs:168) — are
 * built on the device by the first HOST call that needs them (0.3 s for the default 8.9 GB Ed25519
 * table) or explicitly by ecb_warm(). */
int ecb_init(const int* device_ids, int n_dev, ecb_ctx** out);
/* Prepare every device of the context for `op` on batches of up to max_n elements: builds the generator
 * comb the op uses (for the current options) and sizes its work buffers by running the op once on an
 * all-zero batch.  Required before *_dev calls (which never allocate), optional before host calls.
 * op: "ed25519_mul_base", "ed25519_mul", "x25519", "x25519_base", "x448", "ed25519_verify_prehashed",
 * "ed25519_public_from_seed", "ed25519_sign", "bls12_381_g1_from_compressed", and with a curve_id
 * "wei_mul", "wei_mul_base", "wei_decompress", "ecdsa_verify_hashed", "ecdsa_sign_hashed"; the *_vartime forms of
 * the key-generation and signing ops are named with that suffix ("ed25519_sign_vartime", ...). */
int ecb_warm(ecb_ctx* ctx, const char* op, int curve_id, size_t max_n);
/* What device dev_index holds now: "ed25519_comb_w", "ed25519_comb_windows", "<curve>_comb_w",
 * "<curve>_comb_windows" (0 = not built yet), "sm_count". */
int ecb_get_info(ecb_ctx* ctx, int dev_index, const char* key, long* value);
void ecb_destroy(ecb_ctx* ctx);
const char* ecb_last_error(ecb_ctx* ctx);
int ecb_device_count(ecb_ctx* ctx);
/* Tunables (set before first use; unknown keys and out-of-range values give ECB_ERR_INVALID_ARG):
 *   "ed25519_comb_w"        window width of the Ed25519 fixed-base comb, 4..26; 0 (default) = by free device memory, at
 *                           most 24 (11 windows, 8.9 GB).  26 (10 windows, 32 GB + 43 GB while it is built) is opt-in.
 *   "p256r1_comb_w" / "p384r1_comb_w" / "bls12_381_g1_comb_w" / "p256k1_comb_w"
 *                           generator combs of the Weierstrass curves, 4..24; 0 = by free device memory
 *   "chunk"                 elements per pipeline chunk of the host entry points
 *   "ramp"                  halvings of the chunk size at both ends of a batch, 0..4
 *   "inv_per_thread"        batch-inversion chain length;  "inv_fill_per_sm": minimum inversion threads per SM
 *   "inv_block"             batch inversion: 0 one inversion per thread, 2 one per block, 1 (default) per field as measured
 *   "inv_hi"                1 (default): batch inversions of pipelined chunks run on a high-priority side stream
 *   "dev_split"             1: split large device-resident batches over the slot streams (default 0)
 *   "ed25519_entry_stride"  32-bit words between Ed25519 comb entries: 32 (default, one entry per 128-byte line: 1.2x the minimum DRAM traffic) or 24 (packed, 96 B: 25 % smaller table, 1.8x the traffic, same speed)
 *   "ed25519_fused"         small-batch kernel (lanes share a scalar, affine conversion in the same launch): 0 never,
 *                           1 (default) when the batch fits one wave of the device, 2 always
 *   "ed25519_lanes"         lanes per scalar in that kernel: 0 (default) by batch size, or 1, 2, 4, 8
 *   "bls12_381_g1_glv"      1: ecb_wei_mul / ecb_wei_mul_dev on ECB_CURVE_BLS12_381_G1 take every input point to be in the
 *                           prime-order subgroup G1 (e.g. outputs of ecb_bls12_381_g1_from_compressed with check_subgroup = 1,
 *                           or of ecb_wei_mul_base) and split the scalar over the endomorphism (beta x, y) = [-x^2] P of the
 *                           reference's subgroup test (g1.rs:105): half the doublings, 1.4x.  For points of E(Fp) OUTSIDE G1 —
 *                           which Point::mul accepts — the result is then NOT k P; default 0 (bit-exact for every curve point).
 *   "trace"                 1: the fused kernel records per-block phase timestamps (ecb_debug_fused_trace)
 *   "profile"               1: record CUDA events around the kernels of every call on the launching stream,
 *                           read back with ecb_profile_collect */
int ecb_set_option(ecb_ctx* ctx, const char* key, long value);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
unsigned long long ecb_launch_count(ecb_ctx* ctx);

void* ecb_alloc_pinned(size_t bytes);
void ecb_free_pinned(void* p);

/* ---- edwards25519 ---------------------------------------------------------------------- */

/* Point::mul_base(&Scalar) -> to_affine   (src/curve/curve25519.rs:840, :663)
 * k_le: n x 32 B canonical scalars (< l); xy_le: n x 64 B affine x || y, canonical. */
int ecb_ed25519_mul_base(ecb_ctx* ctx, const uint8_t* k_le, size_t n, uint8_t* xy_le, size_t* bad_index);
/* same, output = encode_point(mul_base(k)) (src/protocol/ed25519.rs:27), n x 32 B */
int ecb_ed25519_mul_base_compressed(ecb_ctx* ctx, const uint8_t* k_le, size_t n, uint8_t* enc, size_t* bad_index);
/* &Point * &Scalar -> to_affine   (src/curve/curve25519.rs:1274, :760)
 * xy_le_in: n x 64 B affine points (must satisfy Point::from_coordinate). */
int ecb_ed25519_mul(ecb_ctx* ctx, const uint8_t* k_le, const uint8_t* xy_le_in, size_t n, uint8_t* xy_le_out,
                    size_t* bad_index);
/* ed25519::decode_point / Point::decompress (src/protocol/ed25519.rs:38-59, src/curve/curve25519.rs:772), batched:
 * enc: n x 32 B (y little-endian, sign of x in bit 255); xy_le: n x 64 B canonical affine x || y; ok[i] = 0 and zero
 * bytes where the reference returns None (y >= p, x^2 not a square, x = 0 with the sign bit set).  The output feeds
 * ecb_ed25519_mul directly. */
int ecb_ed25519_decompress(ecb_ctx* ctx, const uint8_t* enc, size_t n, uint8_t* xy_le, uint8_t* ok);
/* ristretto255 (src/curve/curve25519/ristretto255.rs; RFC 9496).  A RistrettoPoint wraps an edwards25519 point and is
 * observable only through its 32-byte encoding, so the batch forms work on encodings:
 *   decompress   RistrettoPoint::decompress (:105): enc n x 32 B -> an Edwards representative, affine x || y (n x 64 B),
 *                ok[i] = 0 and zero bytes for the encodings RFC 9496 rejects (non-canonical or negative s, non-square, ...)
 *   compress     RistrettoPoint::compress (:73) of affine Edwards points: equal group elements give equal bytes
 *   mul          RistrettoPoint::scale (:152): enc_out[i] = compress(k_i * decompress(enc_in[i])); an invalid encoding
 *                fails the call with ECB_ERR_POINT_NOT_ON_CURVE and its index, a scalar >= l with ECB_ERR_NONCANONICAL_SCALAR
 *   mul_base     RistrettoPoint::mul_base (:157): compress(k_i * B) */
int ecb_ristretto255_decompress(ecb_ctx* ctx, const uint8_t* enc, size_t n, uint8_t* xy_le, uint8_t* ok);
int ecb_ristretto255_compress(ecb_ctx* ctx, const uint8_t* xy_le, size_t n, uint8_t* enc);
int ecb_ristretto255_mul(ecb_ctx* ctx, const uint8_t* k_le, const uint8_t* enc_in, size_t n, uint8_t* enc_out, size_t* bad_index);
int ecb_ristretto255_mul_base(ecb_ctx* ctx, const uint8_t* k_le, size_t n, uint8_t* enc_out, size_t* bad_index);
/* protocol::ed25519 verify with k = SHA-512(R||A||M) mod l computed by the caller
 * (src/protocol/ed25519.rs:119-147).  a_enc, r_enc, s_le, k_le: n x 32 B; ok: n x 1 B (0/1). */
int ecb_ed25519_verify_prehashed(ecb_ctx* ctx, const uint8_t* a_enc, const uint8_t* r_enc, const uint8_t* s_le,
                                 const uint8_t* k_le, size_t n, uint8_t* ok);

/* ed25519::PublicKey::verify(msg, sig)   (src/protocol/ed25519.rs:119-147, :200) on raw messages:
 * k = reduce_wide_le(SHA-512(R || A || M)) (:21, :139) is computed on the device.
 * a_enc: n x 32 B public keys; sig: n x 64 B R || S; message i = msgs[msg_off[i] .. msg_off[i+1])
 * (msg_off: n + 1 non-decreasing byte offsets); ok: n x 1 B. */
int ecb_ed25519_verify(ecb_ctx* ctx, const uint8_t* a_enc, const uint8_t* msgs, const uint64_t* msg_off, const uint8_t* sig,
                       size_t n, uint8_t* ok);

/* ---- X25519 / X448 --------------------------------------------------------------------- */

/* ---- Ed25519 key generation and signing -------------------------------------------------
 * SecretKey::public_key (src/protocol/ed25519.rs:81 public_from_seed, :61 expand_secret): seeds n x 32 bytes ->
 * pub n x 32 bytes = encode_point([clamp(SHA-512(seed)[0..32]) mod l] B).
 * Keypair::sign / sign_with_public (ed25519.rs:94-110; SecretKey::sign :112 when pub is NULL: A is derived on
 * the device first): r = SHA-512(prefix || M) mod l, R = encode_point(r B), k = SHA-512(R || A || M) mod l,
 * S = r + k a mod l; sig: n x 64 bytes R || S.  Messages as in ecb_ed25519_verify (concatenated, n + 1 offsets).
 *
 * The entry points under the reference's names are CONSTANT-TIME with respect to the seeds, as the reference is
 * (select_from_table scans every entry, curve25519.rs:862-869; Fermat inverses :155-200): a B and r B come from
 * csrc/ct.cuh — a 48 KB comb in shared memory, every window reads all 8 entries and keeps one by masks, the affine
 * conversion is Montgomery's trick around the fixed Fermat chain — and the work buffers that held key material
 * are cleared on the stream.  The *_vartime entry points are the fast forms (the large comb indexed by the
 * scalar's digits, variable-time safegcd; ~7x the throughput): same bytes, for callers whose GPU is not shared
 * with an adversary.  ecb_ed25519_mul_base_ct is Point::mul_base itself for secret scalars. */
int ecb_ed25519_public_from_seed(ecb_ctx* ctx, const uint8_t* seeds, size_t n, uint8_t* pub);
int ecb_ed25519_sign(ecb_ctx* ctx, const uint8_t* seeds, const uint8_t* pub, const uint8_t* msgs, const uint64_t* msg_off, size_t n,
                     uint8_t* sig);
int ecb_ed25519_public_from_seed_vartime(ecb_ctx* ctx, const uint8_t* seeds, size_t n, uint8_t* pub);
int ecb_ed25519_sign_vartime(ecb_ctx* ctx, const uint8_t* seeds, const uint8_t* pub, const uint8_t* msgs, const uint64_t* msg_off, size_t n,
                             uint8_t* sig);
int ecb_ed25519_mul_base_ct(ecb_ctx* ctx, const uint8_t* k_le, size_t n, uint8_t* xy_le, size_t* bad_index);

/* protocol::x25519::x25519(scalar, u)   (src/protocol/x25519.rs:36): clamps inside, masks bit 255
 * of u, accepts non-canonical u, returns 0 for low-order inputs.  k, u, out: n x 32 B. */
int ecb_x25519(ecb_ctx* ctx, const uint8_t* k, const uint8_t* u, size_t n, uint8_t* out);
/* protocol::x25519::x25519_base(scalar) = x25519(scalar, 9)   (src/protocol/x25519.rs:49; SecretKey::public_key
 * :75): public-key generation.  Same bytes as ecb_x25519 with u = 9, computed with the fixed-base
 * comb instead of the ladder.  k, out: n x 32 B. */
int ecb_x25519_base(ecb_ctx* ctx, const uint8_t* k, size_t n, uint8_t* out);
/* protocol::x448::x448(scalar, u)   (src/protocol/x448.rs:34).  k, u, out: n x 56 B. */
int ecb_x448(ecb_ctx* ctx, const uint8_t* k, const uint8_t* u, size_t n, uint8_t* out);

/* ---- short Weierstrass: p256r1, p384r1, bls12_381 G1 ------------------------------------ */

/* &Point * &Scalar -> to_affine   (src/curve/fiat/curve_macros.rs:321 -> projective.rs:871/:842)
 * k_be: n x SB canonical scalars; xy_be: n x 2FB affine x || y; inf_in: optional n x 1 B, non-zero
 * marks an identity input (its xy is ignored); out_xy_be: n x 2FB (zeros for the identity);
 * out_inf: n x 1 B, 1 when the result is the identity (to_affine() == None, projective.rs:666). */
int ecb_wei_mul(ecb_ctx* ctx, int curve_id, const uint8_t* k_be, const uint8_t* xy_be, const uint8_t* inf_in, size_t n,
                uint8_t* out_xy_be, uint8_t* out_inf, size_t* bad_index);
/* Point::mul_base(&Scalar) -> to_affine   (fiat/curve_macros.rs:55 -> projective.rs:965/:945) */
int ecb_wei_mul_base(ecb_ctx* ctx, int curve_id, const uint8_t* k_be, size_t n, uint8_t* out_xy_be, uint8_t* out_inf,
                     size_t* bad_index);
/* Multi-scalar multiplication: sum_i k_i * P_i -> to_affine (the reference lists it as a wish, TODO.md:48, :99-100; the
 * value is what folding its own `&P * &k` (projective.rs:842) with `+` (:268) over the batch gives).  Bucket method
 * (csrc/msm.cuh): signed windows of up to 16 bits, counting sort of the digits, bucket sums balanced over equal
 * segments of the sorted array (the running time does not depend on how the digits are distributed), running sums
 * over the buckets.  curve_id:
 * ECB_CURVE_BLS12_381_G1 or ECB_CURVE_P256K1.  k_be: n x SB canonical scalars; xy_be: n x 2FB affine points on the curve
 * (any subgroup); out_xy_be: 2FB bytes (zeros for the identity), out_inf: 1 byte.  n == 0 gives the identity.  The
 * batch is sliced over the devices of the context; each reduces its slice and the partial sums are added on the
 * first device — the one operation of this library with a cross-device step. */
int ecb_wei_msm(ecb_ctx* ctx, int curve_id, const uint8_t* k_be, const uint8_t* xy_be, size_t n, uint8_t* out_xy_be, uint8_t* out_inf,
                size_t* bad_index);
/* ecdsa::sign_hashed::<O>(&secret, &nonce, hashed) -> CtOption<Signature> (src/protocol/ecdsa.rs:165-184), batched
 * (SURVEY 8 f.3).  d_be (secret), k_be (nonce), z_be (message scalar): n x SB bytes each;
 * rs_be: n x 2SB bytes r || s; ok[i] = 0 (and zero output) where the reference reports no signature: secret or
 * nonce zero, r = 0 or s = 0 - and for a non-canonical scalar (Scalar::from_bytes -> None).
 * Under the reference's names the secret-dependent steps are CONSTANT-TIME, as the reference's are: k G by masked
 * scans of a shared-memory comb with the complete projective addition (csrc/ct.cuh; projective.rs:340-423, :427),
 * k^-1 by the fixed Fermat chain, work buffers cleared afterwards.  The *_vartime forms are ~6x faster (large comb
 * indexed by the nonce's digits, Jacobian additions that branch on exceptional cases, safegcd): same bytes. */
int ecb_ecdsa_sign_hashed(ecb_ctx* ctx, int curve_id, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* z_be, size_t n,
                          uint8_t* rs_be, uint8_t* ok);
int ecb_ecdsa_sign_hashed_vartime(ecb_ctx* ctx, int curve_id, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* z_be, size_t n,
                                  uint8_t* rs_be, uint8_t* ok);
/* ecdsa::sign::<O>(&secret, &nonce, message) (src/protocol/ecdsa.rs:192): z = O::hash_to_scalar(message) on the device
 * (hash = 256 / 384 / 512, as ecb_ecdsa_verify), then sign_hashed.  Messages concatenated, n + 1 offsets. */
int ecb_ecdsa_sign(ecb_ctx* ctx, int curve_id, int hash, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* msgs,
                   const uint64_t* msg_off, size_t n, uint8_t* rs_be, uint8_t* ok);
int ecb_ecdsa_sign_vartime(ecb_ctx* ctx, int curve_id, int hash, const uint8_t* d_be, const uint8_t* k_be, const uint8_t* msgs,
                   const uint64_t* msg_off, size_t n, uint8_t* rs_be, uint8_t* ok);

/* ---- wire formats either side of the Weierstrass path --------------------------------------------------
 * PointAffine::decompress(&FieldElement, Sign) -> CtOption<PointAffine> (src/curve/affine.rs:48,
 * src/curve/fiat/curve_macros.rs:221; square roots src/curve/sec2/p256r1.rs:68, p384r1.rs:71,
 * src/curve/bls12_381/fp.rs:64), batched.  x_be: n x FB bytes; sign: n bytes, 0 = Sign::Positive (the
 * root whose canonical value is even), non-zero = Sign::Negative (odd) (field_macros.rs:557).
 * out_xy_be: n x 2FB bytes x || y; ok[i] = 1 when the reference returns a present value, else 0 and
 * the output bytes are zero: x >= p (FieldElement::from_bytes -> None, field_macros.rs:15) or
 * x^3 + a x + b not a square. */
int ecb_wei_decompress(ecb_ctx* ctx, int curve_id, const uint8_t* x_be, const uint8_t* sign, size_t n, uint8_t* out_xy_be,
                       uint8_t* ok);
/* BLS12-381 G1 standard (zcash / IETF) encodings, src/curve/bls12_381/serialize.rs.
 * from_compressed: PointAffine::from_compressed (serialize.rs:286; check_subgroup != 0, the
 * prime-order-subgroup test is PointAffine::is_in_subgroup, g1.rs:105) or from_compressed_oncurve_only
 * (serialize.rs:310; check_subgroup == 0).  enc: n x 48 bytes; out_xy_be: n x 96 bytes; ok[i] = 0 (None,
 * output zero) for a clear compression flag, the identity (the affine type cannot hold it), flag
 * misuse, x >= p, x^3 + 4 not a square, and - when checking - points outside the subgroup.
 * to_compressed: Point::to_compressed (serialize.rs:400): x with flag bits 7 (compressed) and 5 (y is the
 * larger root, y > (p-1)/2); inf (n bytes, may be NULL) marks identities, encoded 0xc0 00 .. 00. */
int ecb_bls12_381_g1_from_compressed(ecb_ctx* ctx, const uint8_t* enc, size_t n, int check_subgroup, uint8_t* out_xy_be,
                                     uint8_t* ok);
int ecb_bls12_381_g1_to_compressed(ecb_ctx* ctx, const uint8_t* xy_be, const uint8_t* inf, size_t n, uint8_t* enc);
/* The uncompressed flavour (serialize.rs:22-28, :269, :330-380): 96 bytes x || y big-endian with the same flag bits.
 * from_uncompressed = PointAffine::from_uncompressed (check_subgroup != 0) / from_uncompressed_oncurve_only (0): the
 * compression and sort bits must be clear, both coordinates canonical, the point on the curve (and in G1);
 * ok[i] = 0 and zero output where the reference gives None — which includes the identity encoding (0x40 then zeros),
 * since PointAffine cannot hold it; out_inf (optional, n bytes) is 1 exactly for that encoding.
 * to_uncompressed = Point::to_uncompressed: x || y as given, the identity (inf[i] != 0) as 0x40 then zeros. */
int ecb_bls12_381_g1_from_uncompressed(ecb_ctx* ctx, const uint8_t* enc, size_t n, int check_subgroup, uint8_t* out_xy_be,
                                       uint8_t* out_inf, uint8_t* ok);
int ecb_bls12_381_g1_to_uncompressed(ecb_ctx* ctx, const uint8_t* xy_be, const uint8_t* inf, size_t n, uint8_t* enc);

/* ecdsa::verify_hashed (src/protocol/ecdsa.rs:205-222) after Signature::from_bytes (:399).
 * q_xy_be: n x 2FB public keys; z_be: n x SB message scalars (any value, reduced mod n as
 * digest_to_scalar :340 does); rs_be: n x 2SB r || s (zero or >= n => ok = 0); ok: n x 1 B. */
int ecb_ecdsa_verify_hashed(ecb_ctx* ctx, int curve_id, const uint8_t* q_xy_be, const uint8_t* z_be,
                            const uint8_t* rs_be, size_t n, uint8_t* ok, size_t* bad_index);

/* ecdsa::verify::<O>(public, message, signature)   (src/protocol/ecdsa.rs:228; O::hash_to_scalar :288):
 * z = digest_to_scalar(SHA-2(message)) is computed on the device.  hash = 256, 384 or 512 (the
 * P256R1_Sha256 / _Sha384 / _Sha512 and P384R1_* marker types); messages as in ecb_ed25519_verify. */
int ecb_ecdsa_verify(ecb_ctx* ctx, int curve_id, int hash, const uint8_t* q_xy_be, const uint8_t* msgs, const uint64_t* msg_off,
                     const uint8_t* rs_be, size_t n, uint8_t* ok, size_t* bad_index);

/* ---- device-resident variants (inputs/outputs already in HBM of device `dev_index`) ------ */
int ecb_ed25519_mul_base_dev(ecb_ctx* ctx, int dev_index, const void* d_k_le, size_t n, void* d_xy_le, void* stream);
int ecb_ed25519_mul_dev(ecb_ctx* ctx, int dev_index, const void* d_k_le, const void* d_xy_in, size_t n, void* d_xy_out,
                        void* stream);
int ecb_x25519_base_dev(ecb_ctx* ctx, int dev_index, const void* d_k, size_t n, void* d_out, void* stream);
int ecb_x25519_dev(ecb_ctx* ctx, int dev_index, const void* d_k, const void* d_u, size_t n, void* d_out, void* stream);
int ecb_wei_mul_dev(ecb_ctx* ctx, int dev_index, int curve_id, const void* d_k_be, const void* d_xy_be, size_t n,
                    void* d_out_xy_be, void* d_out_inf, void* stream);
int ecb_wei_mul_base_dev(ecb_ctx* ctx, int dev_index, int curve_id, const void* d_k_be, size_t n, void* d_out_xy_be,
                         void* d_out_inf, void* stream);
int ecb_wei_decompress_dev(ecb_ctx* ctx, int dev_index, int curve_id, const void* d_x_be, const void* d_sign, size_t n,
                           void* d_out_xy_be, void* d_ok, void* stream);
int ecb_bls12_381_g1_from_compressed_dev(ecb_ctx* ctx, int dev_index, const void* d_enc, size_t n, int check_subgroup,
                                         void* d_out_xy_be, void* d_ok, void* stream);
int ecb_ecdsa_sign_hashed_dev(ecb_ctx* ctx, int dev_index, int curve_id, const void* d_d_be, const void* d_k_be, const void* d_z_be,
                              size_t n, void* d_rs_be, void* d_ok, void* stream);
int ecb_ecdsa_sign_hashed_vartime_dev(ecb_ctx* ctx, int dev_index, int curve_id, const void* d_d_be, const void* d_k_be, const void* d_z_be,
                              size_t n, void* d_rs_be, void* d_ok, void* stream);
int ecb_ed25519_public_from_seed_dev(ecb_ctx* ctx, int dev_index, const void* d_seeds, size_t n, void* d_pub, void* stream);
int ecb_ed25519_public_from_seed_vartime_dev(ecb_ctx* ctx, int dev_index, const void* d_seeds, size_t n, void* d_pub, void* stream);
int ecb_ed25519_sign_vartime_dev(ecb_ctx* ctx, int dev_index, const void* d_seeds, const void* d_pub, const void* d_msgs, const void* d_msg_off,
                                 size_t n, void* d_sig, void* stream);
int ecb_ed25519_sign_dev(ecb_ctx* ctx, int dev_index, const void* d_seeds, const void* d_pub, const void* d_msgs, const void* d_msg_off,
                         size_t n, void* d_sig, void* stream);
int ecb_x448_dev(ecb_ctx* ctx, int dev_index, const void* d_k, const void* d_u, size_t n, void* d_out, void* stream);
int ecb_ed25519_verify_prehashed_dev(ecb_ctx* ctx, int dev_index, const void* d_a_enc, const void* d_r_enc, const void* d_s_le,
                                     const void* d_k_le, size_t n, void* d_ok, void* stream);
int ecb_ecdsa_verify_hashed_dev(ecb_ctx* ctx, int dev_index, int curve_id, const void* d_q_xy_be, const void* d_z_be,
                                const void* d_rs_be, size_t n, void* d_ok, void* stream);
/* after a *_dev call and a stream sync: 0, or the error of the first invalid element */
int ecb_dev_status(ecb_ctx* ctx, int dev_index, size_t* bad_index);

/* ---- measurement helpers ----------------------------------------------------------------- */
/* Integer-pipe peak probe.  variant: 0 = IMAD (32-bit), 1 = IMAD.WIDE.U32, 2 = IMAD.WIDE.U32.X carry
 * chains (the field kernels' instruction), 3 = IMAD.HI.U32, 4 = DFMA, 5 = integer add/logic.  Returns multiply-accumulates per second
 * on device `dev_index` and the kernel time. */
int ecb_imad_probe(ecb_ctx* ctx, int dev_index, int variant, int iters, double* macs_per_s, double* ms);
/* Latency probe: cycles per operation of a DEPENDENT chain in one warp per SM (what a small batch pays) and the
 * SM clock the probe ran at.  variant: 0 field mul, 1 field square (GF(2^255-19)), 2 safegcd inversion, 3 Fermat
 * inversion, 4 block-cooperative inversion of `threads` elements (fused.cuh), 5 one 8-word shuffle, 6 mixed
 * extended addition, 7 complete extended addition.  threads: block size, multiple of 32, <= 512. */
int ecb_latency_probe(ecb_ctx* ctx, int dev_index, int variant, int threads, int reps, double* cycles, double* sm_mhz);
/* Field-multiplication throughput on the two multiplier pipes: every thread runs a dependent chain of `reps`
 * GF(2^255-19) products; warps with (warp % fp64_den) < fp64_num use the FP64-pipe field (csrc/fe43.cuh), the
 * others the integer field (csrc/fe25519.cuh); blocks_per_sm blocks of 128 threads per SM.  Returns products per
 * second over the chip and a checksum word (the same for every split: both fields compute the same values). */
int ecb_fieldmul_probe(ecb_ctx* ctx, int dev_index, int fp64_num, int fp64_den, int blocks_per_sm, int reps, double* muls_per_s,
                       double* check);
/* with option "profile" = 1: sum over the calls since the last collect of the device time (ms) of
 * the scalar-multiplication kernel(s) and of the batch-inversion / encoding kernel; synchronises. */
int ecb_profile_collect(ecb_ctx* ctx, int dev_index, double* main_ms, double* finish_ms, int* calls);
/* debug (no device needed): the pipeline chunk schedule the host entry points use for one device's
 * slice [lo, hi) — writes the chunk boundaries (first = lo, last = hi) and returns how many */
long ecb_debug_chunk_plan(size_t lo, size_t hi, size_t chunk, long ramp, size_t* bounds, size_t cap);
/* debug (option "trace" = 1): per block of the last fused small-batch launch four %globaltimer values in ns —
 * start, comb done, inversion done, end.  Synchronises the device; returns the number of blocks recorded. */
long ecb_debug_fused_trace(ecb_ctx* ctx, int dev_index, unsigned long long* out, size_t cap_blocks);
/* debug: copy the device's Ed25519 comb table (niels entries, 96 B each) to host; returns entries */
long ecb_debug_ed25519_table(ecb_ctx* ctx, int dev_index, uint8_t* out, size_t cap_bytes, int* w, int* nwin);

#ifdef __cplusplus
}
#endif
#endif
