// Launch wrappers for the sm_100a kernels (definitions in et_hist.cu, et_pack.cu, et_unpack.cu).
#pragma once
#include <cuda_runtime.h>

#include "et_internal.h"

namespace et {

// ---------------------------------------------------------------- K1
constexpr int kHistThreads = 512;
// counts (device, 256 x u64) must be zeroed by the caller; the kernel adds into it.
cudaError_t launch_histogram(const uint8_t *d_in, size_t n, unsigned long long *d_counts, int num_sms,
                             cudaStream_t stream);

// ---------------------------------------------------------------- K2
constexpr int kPackThreads = 256;
constexpr int kPackItems = 16;  // symbols per thread per tile (one 16-byte load)
constexpr int kPackTileSyms = kPackThreads * kPackItems;

struct PackGeometry {
    const uint8_t *in_aligned;  // d_in rounded down to 16 bytes
    uint32_t misalign;          // d_in - in_aligned
    uint64_t v_end;             // misalign + n (virtual end)
    uint32_t num_tiles;
};
PackGeometry pack_geometry(const void *d_in, size_t n);

// Device scratch the pack kernels need for `num_tiles` tiles.
struct PackScratch {
    unsigned long long *tile_state;  // [num_tiles] decoupled-lookback descriptors
    uint8_t *seam_head;              // [num_tiles]
    uint8_t *seam_tail;              // [num_tiles]
    uint32_t *ticket;                // [1]
};
size_t pack_scratch_bytes(uint32_t num_tiles);
PackScratch pack_scratch_carve(void *base, uint32_t num_tiles);

// d_tables: narrow -> 256 x u32; wide -> 256 x u64 codes followed by 256 x u8 lengths.
cudaError_t launch_pack(const PackGeometry &g, const void *d_tables, bool wide, uint8_t *d_out, uint32_t bit_phase,
                        const PackScratch &s, void *scratch_base, size_t scratch_bytes, int num_sms,
                        cudaStream_t stream, int *launches);

// ---------------------------------------------------------------- K3-K5 (self-synchronising decoder)
constexpr int kUnpackThreads = 256;  // subsequences per tile, warm-up included
constexpr int kUnpackWarm = 8;       // leading subsequences re-decoded from the previous tile
constexpr int kUnpackOwned = kUnpackThreads - kUnpackWarm;
constexpr int kSubseqBits = 128;

struct UnpackGeometry {
    const uint8_t *body_aligned;  // body pointer rounded down to 16 bytes
    uint64_t first_bit;           // bit offset of the first codeword inside body_aligned
    uint64_t end_bit;             // first bit past the stream inside body_aligned
    uint32_t num_tiles;
};
UnpackGeometry unpack_geometry(const void *d_body, size_t body_bytes);

struct UnpackScratch {
    unsigned long long *tile_state;  // [num_tiles]
    uint32_t *ticket;                // [1]
    uint32_t *error_flags;           // [1] bit0 seam mismatch, bit1 invalid code, bit2 no convergence
    unsigned long long *total;       // [1] symbols found in the stream
};
size_t unpack_scratch_bytes(uint32_t num_tiles);
UnpackScratch unpack_scratch_carve(void *base, uint32_t num_tiles);

constexpr uint32_t kErrSeam = 1u, kErrInvalidCode = 2u, kErrNoConvergence = 4u;

// d_lut: kLutSize x u32, d_nodes: trie.  Writes min(total, max_symbols) bytes to d_out.
cudaError_t launch_unpack(const UnpackGeometry &g, const uint32_t *d_lut, const uint32_t *d_nodes, uint8_t *d_out,
                          uint64_t max_symbols, const UnpackScratch &s, void *scratch_base, size_t scratch_bytes,
                          int num_sms, cudaStream_t stream, int *launches);

// ---------------------------------------------------------------- chunked decoder (any prefix code)
constexpr int kChunkBytes = 1024;   // stream bytes per thread
constexpr int kChunkThreads = 128;
size_t chunked_scratch_bytes(uint64_t end_bit);
// Same result contract as launch_unpack; blocks on the stream between fixpoint rounds
// (h_flag: pinned host word).  *rounds_out = sync launches it took.
cudaError_t launch_unpack_chunked(const UnpackGeometry &g, const uint32_t *d_lut, const uint32_t *d_nodes, uint8_t *d_out,
                                  uint64_t max_symbols, void *scratch_base, size_t scratch_bytes, uint32_t *h_flag,
                                  cudaStream_t stream, int *launches, uint32_t *rounds_out);

// ---------------------------------------------------------------- synthetic input generator
cudaError_t launch_synth(uint8_t *d_out, size_t n, uint64_t seed, uint64_t first_index, const uint32_t *d_thresholds,
                         cudaStream_t stream);

}  // namespace et
