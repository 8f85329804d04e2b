// K3-K5 — self-synchronising parallel Huffman decode of an .et body (replaces decode.zig:143-203).
//
// The stream carries no block index, so no thread knows where a codeword starts.  The body
// is cut into 128-bit subsequences, one per thread, held in REGISTERS for the whole kernel:
//   sync   : every thread decodes its subsequence from a guessed start (offset 0) and hands
//            the position at which it ran into the next subsequence to its right neighbour;
//            a thread whose start changed decodes again.  Huffman codes re-synchronise after
//            a few symbols, so this Jacobi iteration reaches its fixpoint in 2-3 rounds.
//            The first kUnpackWarm subsequences of a tile belong to the previous tile and are
//            only there to feed the first owned subsequence a synchronised start.
//   scan   : symbol counts -> block scan -> decoupled look-back across tiles (64-bit).
//   write  : each thread decodes once more from its final start into a shared staging
//            buffer, which is written with aligned 16-byte stores.
// One read of the body, one write of the text: algorithmic HBM bytes only (C + N).
//
// Correctness does not rest on the guess: the look-back descriptor of tile t carries the
// exit position of its last subsequence and tile t+1 compares it with the start it used.
// By induction from tile 0 (true start) "no mismatch" proves every start was the true one;
// any mismatch (or a tile that does not converge) raises a flag and the host reruns the
// stream through the exhaustive-offset path (et_unpack_exhaustive.cu), which has no guess.
#include "et_device.cuh"
#include "et_kernels.cuh"

namespace et {

namespace {

constexpr int kWarps = kUnpackThreads / 32;
constexpr int kStageBytes = 8192;  // output staging per pass; tiles with more symbols loop
constexpr int kMaxRounds = 24;
constexpr uint32_t kExitShift = 56;
constexpr unsigned long long kCountMask = (1ull << kExitShift) - 1;

struct UnpackArgs {
    const uint8_t *body_aligned;
    uint64_t first_bit, end_bit;
    uint64_t byte_lo, byte_hi;  // readable bytes of body_aligned: [byte_lo, byte_hi)
    uint64_t n_subseq;
    uint32_t num_tiles;
    const uint32_t *lut;
    const uint32_t *nodes;
    uint8_t *out;
    uint64_t max_symbols;
    unsigned long long *tile_state;
    uint32_t *ticket;
    uint32_t *error_flags;
    unsigned long long *total;
};

// A code longer than the first-level window: walk the trie with the remaining window bits.
// Returns the code length (symbol in *sym) or 0 when no code matches.
__device__ __noinline__ uint32_t long_code(uint32_t win, uint32_t entry, const uint32_t *__restrict__ nodes,
                                           uint32_t *sym) {
    uint32_t node = entry & 0xFFFFu;
    if (node == kChildNone) return 0;
    for (int b = kLutBits; b < 32; ++b) {
        const uint32_t bit = (win >> (31 - b)) & 1u;
        const uint32_t child = (__ldg(nodes + node) >> (16 * bit)) & 0xFFFFu;
        if (child == kChildNone) return 0;
        if (child & kChildLeaf) {
            *sym = child & 0xFFu;
            return (uint32_t)b + 1u;
        }
        node = child;
    }
    return 0;
}

// Decode the subsequence held in w[0..3] (w[4] = look-ahead word) from bit `start`.
// Counts (and, when WRITE, stores) every symbol that BEGINS before bit 128; returns the
// position of the first codeword at or after bit 128.  TAIL: nothing may end after `lim`.
template <bool TAIL, bool WRITE>
__device__ __forceinline__ uint32_t walk_subseq(const uint32_t (&w)[5], uint32_t start, int lim,
                                                const uint32_t *__restrict__ lut,
                                                const uint32_t *__restrict__ nodes, uint32_t *count,
                                                uint8_t *stage, uint32_t out_idx, uint32_t out_len, bool *bad) {
    uint32_t pos = start, n = 0;
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
        const uint32_t hi = w[wi], lo = w[wi + 1];
        const uint32_t bound = 32u * (wi + 1);
        while (pos < bound) {
            if (TAIL && (int)pos >= lim) { pos = 160; break; }
            const uint32_t win = __funnelshift_l(lo, hi, pos & 31u);  // 32 stream bits from pos
            const uint32_t e = lut[win >> (32 - kLutBits)];
            uint32_t len0 = (e >> 16) & 15u;
            uint32_t sym0 = e & 0xFFu;
            if (len0 == 0) {  // rare: longer than the window, or not a code at all
                len0 = long_code(win, e, nodes, &sym0);
                if (len0 == 0) {
                    *bad = true;
                    pos += 1;
                    continue;
                }
                if (TAIL && (int)(pos + len0) > lim) { pos = 160; break; }
                if (WRITE) {
                    if (out_idx < out_len) stage[out_idx] = (uint8_t)sym0;
                    ++out_idx;
                }
                pos += len0;
                n += 1;
                continue;
            }
            // several codes per lookup as long as all of them begin before bit 128
            const bool multi = !TAIL && (wi < 3 || pos + kLutBits <= (uint32_t)kSubseqBits);
            if (WRITE) {
                const uint32_t len01 = (e >> 20) & 15u;
                if (TAIL && (int)(pos + len0) > lim) { pos = 160; break; }
                if (out_idx < out_len) stage[out_idx] = (uint8_t)sym0;
                ++out_idx;
                if (multi && len01) {
                    if (out_idx < out_len) stage[out_idx] = (uint8_t)(e >> 8);
                    ++out_idx;
                    pos += len01;
                    n += 2;
                } else {
                    pos += len0;
                    n += 1;
                }
            } else {
                if (multi) {
                    pos += (e >> 24) & 15u;
                    n += e >> 28;
                } else {
                    if (TAIL && (int)(pos + len0) > lim) { pos = 160; break; }
                    pos += len0;
                    n += 1;
                }
            }
        }
    }
    *count = n;
    return pos;
}

__device__ __forceinline__ uint32_t warp_inclusive_scan_u32(uint32_t v, uint32_t lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (uint32_t)d) v += up;
    }
    return v;
}

// Look-back over symbol counts.  *nearest receives the descriptor of tile-1.
__device__ __forceinline__ unsigned long long lookback_symbols(const unsigned long long *state, uint32_t tile,
                                                               uint32_t lane, unsigned long long *nearest) {
    unsigned long long exclusive = 0;
    long long base = (long long)tile - 1;
    bool first_round = true;
    for (;;) {
        const long long idx = base - (long long)lane;
        unsigned long long d;
        uint32_t has_prefix, pending;
        do {
            d = idx >= 0 ? ld_relaxed_u64(state + idx) : kStatusPrefix;
            has_prefix = __ballot_sync(0xffffffffu, (d & kStatusMask) == kStatusPrefix);
            pending = __ballot_sync(0xffffffffu, (d & kStatusMask) == 0);
            if (has_prefix) pending &= (1u << (__ffs((int)has_prefix) - 1)) - 1u;
        } while (pending);
        if (first_round) {
            *nearest = __shfl_sync(0xffffffffu, d, 0);
            first_round = false;
        }
        const uint32_t first = has_prefix ? (uint32_t)__ffs((int)has_prefix) - 1u : 31u;
        unsigned long long v = lane <= first ? (d & kCountMask) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        exclusive += v;
        if (has_prefix) return exclusive;
        base -= 32;
    }
}

__device__ __forceinline__ uint32_t load_word_safe(const UnpackArgs &a, uint64_t byte) {
    // big-endian 32-bit word at `byte` of body_aligned; bytes outside the stream read as 0
    if (byte >= a.byte_lo && byte + 4 <= a.byte_hi)
        return bswap32(*reinterpret_cast<const uint32_t *>(a.body_aligned + byte));
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (byte + k >= a.byte_lo && byte + k < a.byte_hi) v |= (uint32_t)a.body_aligned[byte + k] << (24 - 8 * k);
    return v;
}


// 16 bytes of subsequence g into registers as big-endian words, plus the look-ahead word
// (first word of subsequence g+1).  Contains a __syncthreads.
__device__ __forceinline__ void load_subseq(const UnpackArgs &a, long long g, bool active, uint32_t (&w)[5],
                                            uint32_t *warp_sh) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4 raw = make_uint4(0, 0, 0, 0);
    if (active) {
        const uint64_t byte = (uint64_t)g * 16;
        if (byte >= a.byte_lo && byte + 16 <= a.byte_hi) {
            raw = ld_stream_v4(a.body_aligned + byte);
        } else {
            const long long lo = (long long)a.byte_lo - (long long)byte, hi = (long long)a.byte_hi - (long long)byte;
            if (hi > 0 && lo < 16) raw = ld_partial_v4(a.body_aligned + byte, (int)max(lo, 0ll), (int)min(hi, 16ll));
        }
    }
    w[0] = bswap32(raw.x); w[1] = bswap32(raw.y); w[2] = bswap32(raw.z); w[3] = bswap32(raw.w);
    uint32_t next = __shfl_down_sync(0xffffffffu, w[0], 1);
    __syncthreads();  // warp_sh may still be in use by the previous phase
    if (lane == 0) warp_sh[warp] = w[0];
    __syncthreads();
    if (lane == 31) {
        if (warp + 1 < (uint32_t)kWarps)
            next = warp_sh[warp + 1];
        else
            next = (g + 1 >= 0 && (uint64_t)(g + 1) < a.n_subseq) ? load_word_safe(a, (uint64_t)(g + 1) * 16) : 0u;
    }
    w[4] = next;
}

// Final pass of a tile: every owning thread decodes its subsequence from its resolved start
// into the staging buffer (positions from the block scan), then the block stores the
// staged text with aligned 16-byte writes.  Tiles with more symbols than the staging
// buffer holds go round the loop again.
template <bool TAIL>
__device__ __forceinline__ void write_tile(const UnpackArgs &a, const uint32_t (&w)[5], uint32_t start, int lim,
                                           bool owned, uint32_t my_cnt, uint32_t my_off, uint32_t tile_total,
                                           unsigned long long out_base, const uint32_t *lut_sh, uint8_t *stage,
                                           bool *bad) {
    const uint32_t tid = threadIdx.x;
    for (uint32_t chunk_lo = 0; chunk_lo < tile_total; chunk_lo += kStageBytes) {
        const unsigned long long g0 = out_base + chunk_lo;
        if (g0 >= a.max_symbols) break;
        uint32_t clen = min((uint32_t)kStageBytes, tile_total - chunk_lo);
        if (g0 + clen > a.max_symbols) clen = (uint32_t)(a.max_symbols - g0);
        uint8_t *dst = a.out + g0;
        const uint32_t align = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
        if (owned && my_cnt && my_off < chunk_lo + clen && my_off + my_cnt > chunk_lo) {
            uint32_t dummy;
            walk_subseq<TAIL, true>(w, start, lim, lut_sh, a.nodes, &dummy, stage + align, my_off - chunk_lo, clen, bad);
        }
        __syncthreads();
        uint8_t *gbase = dst - align;  // staging byte k <-> gbase[k]
        const uint32_t s_lo = align, s_hi = align + clen;
        const uint4 *stage4 = reinterpret_cast<const uint4 *>(stage);
        for (uint32_t c = tid; c * 16 < s_hi; c += kUnpackThreads) {
            const uint32_t k0 = c * 16;
            if (k0 >= s_lo && k0 + 16 <= s_hi) {
                st_stream_v4(gbase + k0, stage4[c]);
            } else {
                for (uint32_t k = max(k0, s_lo); k < min(k0 + 16, s_hi); ++k) gbase[k] = stage[k];
            }
        }
        __syncthreads();
    }
}

template <bool TAIL>
__device__ __forceinline__ void unpack_tile(const UnpackArgs &a, uint32_t tile, const uint32_t *lut_sh,
                                            uint8_t *stage, uint32_t *exit_sh, uint32_t *warp_sh,
                                            unsigned long long *base_sh) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long g = (long long)tile * kUnpackOwned - kUnpackWarm + tid;  // subsequence index
    const bool active = g >= 0 && (uint64_t)g < a.n_subseq;
    const bool owned = active && tid >= (uint32_t)kUnpackWarm;

    uint32_t w[5];
    load_subseq(a, g, active, w, warp_sh);
    int lim = 0;
    if (TAIL) {
        const long long l = (long long)a.end_bit - g * (long long)kSubseqBits;
        lim = (int)max(0ll, min(l, 160ll));
    }

    // ---- sync: Jacobi iteration on start positions
    const bool is_stream_head = tile == 0 && tid == (uint32_t)kUnpackWarm;  // g == 0: true start known
    const bool fixed = tid == 0 || is_stream_head;
    uint32_t start = is_stream_head ? (uint32_t)a.first_bit : 0u;
    uint32_t cnt = 0, my_exit = 0;
    bool bad = false;
    if (active) my_exit = walk_subseq<TAIL, false>(w, start, lim, lut_sh, a.nodes, &cnt, nullptr, 0, 0, &bad) - kSubseqBits;
    bool converged = false;
    for (int round = 0; round < kMaxRounds; ++round) {
        __syncthreads();  // previous round's readers are done (also orders warp_sh reuse)
        exit_sh[tid] = my_exit;
        __syncthreads();
        bool changed = false;
        if (active && !fixed) {
            const uint32_t ns = exit_sh[tid - 1];
            if (ns != start) {
                start = ns;
                my_exit = walk_subseq<TAIL, false>(w, start, lim, lut_sh, a.nodes, &cnt, nullptr, 0, 0, &bad) - kSubseqBits;
                changed = true;
            }
        }
        if (!__syncthreads_or(changed)) {
            converged = true;
            break;
        }
    }
    if (!converged && tid == 0) atomicOr(a.error_flags, kErrNoConvergence);

    // ---- scan: symbols owned by this tile, then the tile's place in the output
    const uint32_t my_cnt = owned ? cnt : 0u;
    const uint32_t incl = warp_inclusive_scan_u32(my_cnt, lane);
    __syncthreads();  // warp_sh was read by lane 31s above; exit_sh readers done
    if (lane == 31) warp_sh[warp] = incl;
    if (tid == (uint32_t)kUnpackWarm) exit_sh[0] = start;        // start the first owned subsequence used
    if (tid == kUnpackThreads - 1) exit_sh[1] = my_exit & 63u;   // exit of the last owned subsequence
    __syncthreads();
    uint32_t warp_off = 0, tile_total = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) {
        const uint32_t s = warp_sh[q];
        if (q < (int)warp) warp_off += s;
        tile_total += s;
    }
    const uint32_t my_off = warp_off + incl - my_cnt;
    if (warp == 0) {
        const unsigned long long exit_tag = (unsigned long long)exit_sh[1] << kExitShift;
        unsigned long long before = 0;
        if (tile != 0) {
            if (lane == 0) st_relaxed_u64(a.tile_state + tile, kStatusAggregate | exit_tag | tile_total);
            unsigned long long nearest = 0;
            before = lookback_symbols(a.tile_state, tile, lane, &nearest);
            // the start we synchronised onto must be where the previous tile really ended
            if (lane == 0 && ((nearest >> kExitShift) & 63u) != exit_sh[0]) atomicOr(a.error_flags, kErrSeam);
        }
        if (lane == 0) {
            st_relaxed_u64(a.tile_state + tile, kStatusPrefix | exit_tag | (before + tile_total));
            *base_sh = before;
            if (tile == a.num_tiles - 1) *a.total = before + tile_total;
        }
    }
    __syncthreads();
    const unsigned long long out_base = *base_sh;

    bad = false;  // speculative rounds may legitimately have hit non-codes; only the final path counts
    write_tile<TAIL>(a, w, start, lim, owned, my_cnt, my_off, tile_total, out_base, lut_sh, stage, &bad);
    if (bad && owned) atomicOr(a.error_flags, kErrInvalidCode);
}

__global__ void __launch_bounds__(kUnpackThreads) unpack_kernel(const UnpackArgs a) {
    __shared__ __align__(16) uint32_t lut_sh[kLutSize];
    __shared__ __align__(16) uint8_t stage[kStageBytes + 32];
    __shared__ uint32_t exit_sh[kUnpackThreads];
    __shared__ uint32_t warp_sh[kWarps];
    __shared__ unsigned long long base_sh;
    __shared__ uint32_t tile_sh, abort_sh;

    for (int i = threadIdx.x; i < kLutSize; i += kUnpackThreads) lut_sh[i] = a.lut[i];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            tile_sh = atomicAdd(a.ticket, 1u);
            abort_sh = ld_relaxed_u32(a.error_flags) & (kErrSeam | kErrNoConvergence);
        }
        __syncthreads();
        const uint32_t tile = tile_sh;
        if (tile >= a.num_tiles) break;
        if (abort_sh) {
            // A guess was wrong somewhere: the host will rerun the stream through the chunked
            // path.  Drain the tickets, publishing descriptors so that no look-back waits forever.
            if (threadIdx.x == 0) st_relaxed_u64(a.tile_state + tile, kStatusPrefix);
            continue;
        }
        // last owned subsequence plus its 32-bit look-ahead reaches past the stream end?
        const unsigned long long reach = ((unsigned long long)tile * kUnpackOwned + kUnpackOwned) * kSubseqBits + 32;
        if (reach > a.end_bit)
            unpack_tile<true>(a, tile, lut_sh, stage, exit_sh, warp_sh, &base_sh);
        else
            unpack_tile<false>(a, tile, lut_sh, stage, exit_sh, warp_sh, &base_sh);
    }
}

}  // namespace

UnpackGeometry unpack_geometry(const void *d_body, size_t body_bytes) {
    UnpackGeometry g;
    const uintptr_t p = reinterpret_cast<uintptr_t>(d_body);
    const uint32_t mis = (uint32_t)(p & 15u);
    g.body_aligned = reinterpret_cast<const uint8_t *>(p - mis);
    g.first_bit = (uint64_t)mis * 8;
    g.end_bit = ((uint64_t)mis + body_bytes) * 8;
    const uint64_t n_subseq = (g.end_bit + kSubseqBits - 1) / kSubseqBits;
    g.num_tiles = body_bytes ? (uint32_t)((n_subseq + kUnpackOwned - 1) / kUnpackOwned) : 0u;
    return g;
}

size_t unpack_scratch_bytes(uint32_t num_tiles) { return 32 + (size_t)num_tiles * 8; }
UnpackScratch unpack_scratch_carve(void *base, uint32_t num_tiles) {
    (void)num_tiles;
    UnpackScratch s;
    uint8_t *p = static_cast<uint8_t *>(base);
    s.ticket = reinterpret_cast<uint32_t *>(p);
    s.error_flags = reinterpret_cast<uint32_t *>(p + 4);
    s.total = reinterpret_cast<unsigned long long *>(p + 8);
    s.tile_state = reinterpret_cast<unsigned long long *>(p + 32);
    return s;
}

cudaError_t launch_unpack(const UnpackGeometry &g, const uint32_t *d_lut, const uint32_t *d_nodes, uint8_t *d_out,
                          uint64_t max_symbols, const UnpackScratch &s, void *scratch_base, size_t scratch_bytes,
                          int num_sms, cudaStream_t stream, int *launches) {
    cudaError_t err = cudaMemsetAsync(scratch_base, 0, scratch_bytes, stream);
    if (err != cudaSuccess) return err;
    if (g.num_tiles == 0) return cudaSuccess;
    UnpackArgs a;
    a.body_aligned = g.body_aligned;
    a.first_bit = g.first_bit;
    a.end_bit = g.end_bit;
    a.byte_lo = g.first_bit >> 3;
    a.byte_hi = g.end_bit >> 3;
    a.n_subseq = (g.end_bit + kSubseqBits - 1) / kSubseqBits;
    a.num_tiles = g.num_tiles;
    a.lut = d_lut;
    a.nodes = d_nodes;
    a.out = d_out;
    a.max_symbols = max_symbols;
    a.tile_state = s.tile_state;
    a.ticket = s.ticket;
    a.error_flags = s.error_flags;
    a.total = s.total;
    unsigned grid = (unsigned)num_sms * 6u;  // persistent; ~26 KiB smem and 256 threads per CTA
    if (grid > g.num_tiles) grid = g.num_tiles;
    unpack_kernel<<<grid, kUnpackThreads, 0, stream>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace et
