// mont_kinds.cuh — which primes get a dedicated Montgomery reduction (see MontKind in mont.cuh).
#pragma once
#include "mont.cuh"
#include "params_gen.cuh"

namespace ecb {

template <>
struct MontKind<P256_FP> {
    static constexpr int kind = 1;
};

template <>
struct MontKind<P384_FP> {
    static constexpr int kind = 2;
};

template <>
struct MontKind<K256_FP> {
    static constexpr int kind = 3;
};

}  // namespace ecb
