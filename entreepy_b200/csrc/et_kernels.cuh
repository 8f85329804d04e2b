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
constexpr int kUnpackStageBytes = 12288;  // text of one tile staged in shared memory (avg code >= 2.6 bits)

// Where the stream sits relative to the 16-byte grid the subsequences are cut on.  All bit
// positions are relative to body_aligned.
struct UnpackGeometry {
    const uint8_t *body_aligned;  // 16-byte aligned base
    uint64_t byte_lo, byte_hi;    // readable bytes: [byte_lo, byte_hi)
    uint64_t own_begin_bit;       // symbols that begin in [own_begin_bit, own_end_bit) are decoded
    uint64_t own_end_bit;
    uint64_t end_bit;             // no code may extend past this bit (end of the stream / of the readable range)
    bool head_known;              // head_bit is a true codeword boundary (else the first start is a guess)
    uint64_t head_bit;
    uint32_t num_tiles;
};
UnpackGeometry unpack_geometry(const void *d_body, size_t body_bytes);
// A shard: d_range 16-byte aligned, own_begin_byte a multiple of 16, own_end_byte a multiple of 16
// unless the stream ends there; range_bytes >= own_end_byte + 16 unless the stream ends there.
UnpackGeometry unpack_geometry_shard(const void *d_range, size_t range_bytes, size_t own_begin_byte, size_t own_end_byte,
                                     long long head_bit);

struct UnpackScratch {
    unsigned long long *tile_state;  // [num_tiles]
    uint32_t *ticket;                // [1]
    uint32_t *error_flags;           // [1] bit0 seam mismatch, bit1 invalid code, bit2 no convergence
    unsigned long long *total;       // [1] symbols found
    uint32_t *entry_exit;            // [2] start used by the first owned subsequence / exit of the last one (bits past it)
};
size_t unpack_scratch_bytes(uint32_t num_tiles);
UnpackScratch unpack_scratch_carve(void *base, uint32_t num_tiles);

constexpr uint32_t kErrSeam = 1u, kErrInvalidCode = 2u, kErrNoConvergence = 4u;

// d_clut/d_wlut: kLutSize x u32 each, d_nodes: trie.  Writes min(total, max_symbols) bytes to d_out.
cudaError_t launch_unpack(const UnpackGeometry &g, const uint32_t *d_clut, const uint32_t *d_wlut, const uint32_t *d_nodes,
                          uint8_t *d_out, uint64_t max_symbols, const UnpackScratch &s, void *scratch_base,
                          size_t scratch_bytes, int num_sms, cudaStream_t stream, int *launches);

// ---------------------------------------------------------------- chunked decoder (any prefix code)
constexpr int kChunkBytes = 1024;   // stream bytes per thread
constexpr int kChunkThreads = 128;
size_t chunked_scratch_bytes(const UnpackGeometry &g);
// Same result contract as launch_unpack; blocks on the stream between fixpoint rounds
// (h_flag: pinned host word).  *rounds_out = sync launches it took.
cudaError_t launch_unpack_chunked(const UnpackGeometry &g, const uint32_t *d_clut, const uint32_t *d_wlut,
                                  const uint32_t *d_nodes, uint8_t *d_out,
                                  uint64_t max_symbols, void *scratch_base, size_t scratch_bytes, uint32_t *h_flag,
                                  cudaStream_t stream, int *launches, uint32_t *rounds_out);

// ---------------------------------------------------------------- synthetic input generator
cudaError_t launch_synth(uint8_t *d_out, size_t n, uint64_t seed, uint64_t first_index, const uint32_t *d_thresholds,
                         cudaStream_t stream);

}  // namespace et
