// philox.cuh -- Philox4x32-10 (Salmon et al., SC'11) and the 53-bit uniform recipe shared
// with the CPU checker used by tests/.  Counter-based: the draw for (start, walk, step) does not
// depend on launch geometry, call order or sharding.
#pragma once
#include <stdint.h>

namespace pb200 {

struct Philox4 { uint32_t v[4]; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

// numerator of a 53-bit uniform, numpy legacy random_sample() bit recipe
__host__ __device__ __forceinline__ uint64_t uniform53(uint32_t a, uint32_t b) {
    return ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
}

}  // namespace pb200
