// Small device helpers shared by the kernels.
#pragma once
#include <cstdint>

namespace et {

// Streaming 16-byte load: read-only path, do not keep in L1 (each input byte is used once per pass).
__device__ __forceinline__ uint4 ld_stream_v4(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_v4(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
// Tile descriptors of the decoupled look-back: one 64-bit word, relaxed gpu-scope accesses
// (the word itself is the only payload, so no fence is needed around it).
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// 16 aligned bytes of which only [lo, hi) (byte offsets relative to p) may be touched; the
// rest read as zero.  Used at the ragged ends of a buffer so no byte outside it is loaded.
__device__ __forceinline__ uint4 ld_partial_v4(const uint8_t *p, int lo, int hi) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; ++k)
        if (k >= lo && k < hi) w[k >> 2] |= (uint32_t)p[k] << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr unsigned long long kStatusShift = 62;
constexpr unsigned long long kStatusAggregate = 1ull << kStatusShift;
constexpr unsigned long long kStatusPrefix = 2ull << kStatusShift;
constexpr unsigned long long kStatusMask = 3ull << kStatusShift;

}  // namespace et
