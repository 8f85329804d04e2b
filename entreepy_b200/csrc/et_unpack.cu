// K3-K5 — self-synchronising parallel Huffman decode of an .et body (replaces decode.zig:143-203).
//
// The stream carries no block index, so no thread knows where a codeword starts.  The body
// is cut into 128-bit subsequences, one per thread, held in REGISTERS for the whole kernel:
//   sync   : every thread decodes its subsequence from a guessed start (offset 0) and hands
//            the position at which it ran into the next subsequence to its right neighbour
//            (warp shuffle; shared memory only across warp boundaries); a thread whose start
//            changed decodes again.  Huffman codes re-synchronise after a few symbols, so this
//            Jacobi iteration reaches its fixpoint in 2-3 rounds.  The first kUnpackWarm
//            subsequences of a tile belong to the previous tile and are only there to feed
//            the first owned subsequence a synchronised start.
//   scan   : symbol counts -> block scan -> decoupled look-back across tiles (64-bit).
//   write  : each thread decodes once more from its final start into a shared staging
//            buffer, which is written with aligned 16-byte stores.
// One read of the body, one write of the text: algorithmic HBM bytes only (C + N).
//
// The kernel is instruction-bound, not HBM-bound (ncu: profiles/), so the inner loops are
// written for instruction count: position and symbol count live in ONE register (bits 0-8
// and 9+), every table entry is pre-packed so that a lookup is followed by a single add,
// and the tables are addressed through 32-bit shared-window addresses.  Anything unusual
// (codes longer than the 12-bit window, the ragged end of the stream, a tile that overflows
// the staging buffer) leaves the fast loops through a marker bit and is redone by the
// generic walker.
//
// Correctness does not rest on the guess: the look-back descriptor of tile t carries the
// exit position of its last subsequence and tile t+1 compares it with the start it used.
// By induction from the first tile (true start) "no mismatch" proves every start was the
// true one; any mismatch (or a tile that does not converge) raises a flag and the host
// reruns the stream through the chunked decoder (et_unpack_chunked.cu), which has no guess.
#include "et_device.cuh"
#include "et_kernels.cuh"

namespace et {

namespace {

constexpr int kWarps = kUnpackThreads / 32;
constexpr int kMaxBlockRounds = 12;
constexpr uint32_t kExitShift = 56;
constexpr unsigned long long kCountMask = (1ull << kExitShift) - 1;
constexpr uint32_t kPosMask = 0x1ffu;  // position field of a packed walk state (bit 8 = marker)

struct UnpackArgs {
    const uint8_t *body_aligned;
    uint64_t end_bit;           // nothing may be decoded past this bit (end of the readable stream)
    uint64_t byte_lo, byte_hi;  // readable bytes of body_aligned: [byte_lo, byte_hi)
    long long g_first;          // first owned subsequence
    long long g_own_end;        // one past the last owned subsequence
    long long g_max;            // one past the last subsequence that holds readable bytes
    uint32_t head_known;        // the first owned subsequence starts at head_bit (a true codeword boundary)
    uint32_t head_bit;          // ... relative to that subsequence
    uint32_t num_tiles;
    const uint32_t *clut;
    const uint32_t *wlut;
    const uint32_t *nodes;
    uint8_t *out;
    uint64_t max_symbols;
    unsigned long long *tile_state;
    uint32_t *ticket;
    uint32_t *error_flags;
    unsigned long long *total;
    uint32_t *entry_exit;  // [0] start the first owned subsequence used, [1] exit of the last owned one
};

// ------------------------------------------------------------------ shared-window accessors
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16_hi(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1+2];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u8_1(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0+1], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ------------------------------------------------------------------ fast walkers
// Packed state c: bits 0-8 position inside the subsequence (bit 8 set = "leave the fast path"),
// bits 9+ symbol count (count walk) or staging address (write walk).
// Table index of the 12-bit window at the current position, as a byte offset into a u32 table.
__device__ __forceinline__ uint32_t window_offset(uint32_t hi, uint32_t lo, uint32_t c) {
    return (__funnelshift_l(lo, hi, c) >> (32 - kLutBits - 2)) & ((kLutSize - 1) << 2);
}

// A code longer than the first-level window: walk the trie with the remaining window bits.
// Returns the code length (symbol in *sym) or 0 when no code matches.
__device__ __noinline__ uint32_t long_code(uint32_t win, uint32_t node, const uint32_t *__restrict__ nodes,
                                           uint32_t *sym) {
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

// The fast loops stopped on a marker: the code at the current position is longer than the
// window.  Resolve that one code through the trie and return the add for it (len | 1 << 9),
// or 0 when the bits are no code at all (the caller then gives the subsequence to the generic
// walker).
__device__ __forceinline__ uint32_t long_code_add(uint32_t hi, uint32_t lo, uint32_t c, uint32_t wlut_s,
                                                  const uint32_t *__restrict__ nodes, uint32_t *sym) {
    const uint32_t win = __funnelshift_l(lo, hi, c);
    const uint32_t node = lds_u32(wlut_s + ((win >> (32 - kLutBits - 2)) & ((kLutSize - 1) << 2))) & 0xffffu;
    const uint32_t len = long_code(win, node, nodes, sym);
    return len ? (len | (1u << 9)) : 0u;
}

// Counts every symbol that begins before bit 128 starting from `start`; returns the packed
// state (position = first codeword boundary at or after 128; marker bit set = not a code,
// redo with the generic walker).
// The position is kept relative to the word being decoded (32 is subtracted after each of
// the first three words) so that every loop test is one bit test.
__device__ __forceinline__ uint32_t count_walk_fast(const uint32_t (&w)[5], uint32_t start, uint32_t clut_s,
                                                    uint32_t wlut_s, const uint32_t *__restrict__ nodes) {
    uint32_t c = start;
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
        const uint32_t hi = w[wi], lo = w[wi + 1];
        for (;;) {
            if (wi < 3) {
                while (!(c & 0x1e0u)) c += lds_u16(clut_s + window_offset(hi, lo, c));
            } else {
                // every code of a 12-bit window begins before bit 128 as long as the window does not cross it
                while ((c & kPosMask) <= (uint32_t)(32 - kLutBits)) c += lds_u16(clut_s + window_offset(hi, lo, c));
                while (!(c & 0x1e0u)) c += lds_u16_hi(clut_s + window_offset(hi, lo, c));
            }
            if (!(c & kLutMarker)) break;
            uint32_t sym;
            const uint32_t add = long_code_add(hi, lo, c, wlut_s, nodes, &sym);
            if (!add) return c + 32u * wi;
            c += add - kLutMarker;
        }
        if (wi < 3) c -= 32u;
    }
    return c + 96u;
}

// Decodes from `start` into the staging buffer at shared address o_addr.  A marker entry
// stores one garbage byte inside the thread's own output range (overwritten right after by
// the symbol of the long code) and has no second symbol.
__device__ __forceinline__ uint32_t write_walk_fast(const uint32_t (&w)[5], uint32_t start, uint32_t o_addr,
                                                    uint32_t clut_s, uint32_t wlut_s,
                                                    const uint32_t *__restrict__ nodes) {
    uint32_t c = start | (o_addr << 9);
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
        const uint32_t hi = w[wi], lo = w[wi + 1];
        for (;;) {
            if (wi < 3) {
                while (!(c & 0x1e0u)) {
                    const uint32_t e = lds_u32(wlut_s + window_offset(hi, lo, c));
                    const uint32_t o = c >> 9;
                    sts_u8(o, e);
                    if (e & (2u << 25)) sts_u8_1(o, e >> 8);
                    c += e >> 16;
                }
            } else {
                while ((c & kPosMask) <= (uint32_t)(32 - kLutBits)) {
                    const uint32_t e = lds_u32(wlut_s + window_offset(hi, lo, c));
                    const uint32_t o = c >> 9;
                    sts_u8(o, e);
                    if (e & (2u << 25)) sts_u8_1(o, e >> 8);
                    c += e >> 16;
                }
                while (!(c & 0x1e0u)) {
                    const uint32_t off = window_offset(hi, lo, c);
                    const uint32_t a = lds_u16_hi(clut_s + off);
                    sts_u8(c >> 9, lds_u32(wlut_s + off));
                    c += a;
                }
            }
            if (!(c & kLutMarker)) break;
            uint32_t sym;
            const uint32_t add = long_code_add(hi, lo, c, wlut_s, nodes, &sym);
            if (!add) return c + 32u * wi;
            sts_u8(c >> 9, sym);
            c += add - kLutMarker;
        }
        if (wi < 3) c -= 32u;
    }
    return c + 96u;
}

// ------------------------------------------------------------------ generic walker
// One symbol at a time with every check: codes up to 32 bits, nothing decoded past `lim`
// (bits, relative to the subsequence), staging writes clipped to [0, out_len).  Used for the
// tiles at the end of the stream, for subsequences that hit a marker and for tiles whose text
// does not fit the staging buffer.  Returns the position reached (>= 128, or 160 when the
// stream ended first); bit 31 of *count_bad flags a bit pattern that is no code.
template <bool WRITE>
__device__ __noinline__ uint32_t walk_generic(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4,
                                              uint32_t start, int lim, uint32_t clut_s, uint32_t wlut_s,
                                              const uint32_t *__restrict__ nodes, uint32_t *count_bad, uint8_t *stage,
                                              uint32_t out_idx, uint32_t out_len) {
    const uint32_t w[5] = {w0, w1, w2, w3, w4};
    uint32_t pos = start, n = 0, bad = 0;
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
        const uint32_t hi = w[wi], lo = w[wi + 1];
        const uint32_t bound = 32u * (wi + 1);
        while (pos < bound) {
            if ((int)pos >= lim) { pos = 160; break; }
            const uint32_t win = __funnelshift_l(lo, hi, pos & 31u);  // 32 stream bits from pos
            const uint32_t off = (win >> (32 - kLutBits - 2)) & ((kLutSize - 1) << 2);
            const uint32_t a = lds_u16_hi(clut_s + off);  // len0 | 1 << 9, or marker
            const uint32_t e = lds_u32(wlut_s + off);
            uint32_t len0 = a & 0xffu, sym0 = e & 0xffu;
            if (a & kLutMarker) {  // longer than the window, or not a code at all
                len0 = long_code(win, e & 0xffffu, nodes, &sym0);
                if (len0 == 0) {
                    bad = 0x80000000u;
                    pos += 1;
                    continue;
                }
            }
            if ((int)(pos + len0) > lim) { pos = 160; break; }
            if (WRITE) {
                if (out_idx < out_len) stage[out_idx] = (uint8_t)sym0;
                ++out_idx;
            }
            pos += len0;
            n += 1;
        }
    }
    *count_bad = n | bad;
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

struct TileShared {
    uint32_t exit_warp[kWarps];   // exit of lane 31 of each warp (sync rounds)
    uint32_t first_word[kWarps];  // first stream word of lane 0 of each warp (look-ahead of the warp before)
    uint32_t warp_sum[kWarps];
    uint32_t entry_used, last_exit;
    unsigned long long base;
    uint32_t tile, abort;
};

// 16 bytes of subsequence g into registers as big-endian words, plus the look-ahead word
// (first word of subsequence g+1).  Contains a __syncthreads.
__device__ __forceinline__ void load_subseq(const UnpackArgs &a, long long g, bool active, uint32_t (&w)[5],
                                            TileShared &sh) {
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
    if (lane == 0) sh.first_word[warp] = w[0];
    __syncthreads();
    if (lane == 31) {
        if (warp + 1 < (uint32_t)kWarps)
            next = sh.first_word[warp + 1];
        else
            next = (g + 1 >= 0 && g + 1 < a.g_max) ? load_word_safe(a, (uint64_t)(g + 1) * 16) : 0u;
    }
    w[4] = next;
}

template <bool TAIL>
__device__ __forceinline__ void unpack_tile(const UnpackArgs &a, uint32_t tile, uint32_t clut_s, uint32_t wlut_s,
                                            uint8_t *stage, TileShared &sh) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long g = a.g_first + (long long)tile * kUnpackOwned - kUnpackWarm + tid;  // subsequence index
    const bool active = g >= 0 && g < a.g_max && g * 16 + 16 > (long long)a.byte_lo;
    const bool owned = active && tid >= (uint32_t)kUnpackWarm && g < a.g_own_end;

    uint32_t w[5];
    load_subseq(a, g, active, w, sh);
    int lim = 1 << 30;
    if (TAIL) {
        const long long l = (long long)a.end_bit - g * (long long)kSubseqBits;
        lim = (int)max(0ll, min(l, 160ll));
    }

    auto count_walk = [&](uint32_t start, uint32_t *cnt) -> uint32_t {
        if (!TAIL) {
            const uint32_t c = count_walk_fast(w, start, clut_s, wlut_s, a.nodes);
            if (!(c & kLutMarker)) {
                *cnt = c >> 9;
                return (c & kPosMask) - kSubseqBits;
            }
        }
        uint32_t cb;
        const uint32_t pos = walk_generic<false>(w[0], w[1], w[2], w[3], w[4], start, lim, clut_s, wlut_s, a.nodes, &cb,
                                                 nullptr, 0, 0);
        *cnt = cb & 0x7fffffffu;
        return pos - kSubseqBits;
    };

    // ---- sync: Jacobi iteration on start positions; neighbours inside a warp talk through
    // shuffles, warps through shared memory
    const bool is_head = a.head_known && tile == 0 && tid == (uint32_t)kUnpackWarm;  // true start known
    const bool fixed = !active || tid == 0 || is_head;
    uint32_t start = is_head ? a.head_bit : 0u;
    uint32_t cnt = 0, my_exit = 0;
    if (active) my_exit = count_walk(start, &cnt);
    uint32_t carry_in = 0;  // lane 0: exit of the previous warp's last lane
    bool converged = false;
    for (int round = 0; round < kMaxBlockRounds; ++round) {
        for (int inner = 0; inner < 40; ++inner) {
            uint32_t in = __shfl_up_sync(0xffffffffu, my_exit, 1);
            if (lane == 0) in = carry_in;
            const bool changed = !fixed && in != start && (lane != 0 || round > 0);
            if (!__any_sync(0xffffffffu, changed)) break;
            if (changed) {
                start = in;
                my_exit = count_walk(start, &cnt);
            }
        }
        if (lane == 31) sh.exit_warp[warp] = my_exit;
        __syncthreads();
        bool stale = false;
        if (lane == 0 && warp > 0) {
            carry_in = sh.exit_warp[warp - 1];
            stale = !fixed && carry_in != start;
        }
        if (!__syncthreads_or(stale)) {
            converged = true;
            break;
        }
    }
    if (!converged && tid == 0) atomicOr(a.error_flags, kErrNoConvergence);

    // ---- scan: symbols owned by this tile, then the tile's place in the output
    const uint32_t my_cnt = owned ? cnt : 0u;
    const uint32_t incl = warp_inclusive_scan_u32(my_cnt, lane);
    if (lane == 31) sh.warp_sum[warp] = incl;
    if (tid == (uint32_t)kUnpackWarm) sh.entry_used = start;  // start the first owned subsequence used
    if (owned && (tid == kUnpackThreads - 1 || g + 1 >= a.g_own_end)) sh.last_exit = my_exit & 63u;
    __syncthreads();
    uint32_t warp_off = 0, tile_total = 0;
#pragma unroll
    for (int q = 0; q < kWarps; ++q) {
        const uint32_t s = sh.warp_sum[q];
        if (q < (int)warp) warp_off += s;
        tile_total += s;
    }
    const uint32_t my_off = warp_off + incl - my_cnt;
    if (warp == 0) {
        const unsigned long long exit_tag = (unsigned long long)sh.last_exit << kExitShift;
        unsigned long long before = 0;
        if (tile != 0) {
            if (lane == 0) st_relaxed_u64(a.tile_state + tile, kStatusAggregate | exit_tag | tile_total);
            unsigned long long nearest = 0;
            before = lookback_symbols(a.tile_state, tile, lane, &nearest);
            // the start we synchronised onto must be where the previous tile really ended
            if (lane == 0 && ((nearest >> kExitShift) & 63u) != sh.entry_used) atomicOr(a.error_flags, kErrSeam);
        } else if (lane == 0) {
            a.entry_exit[0] = sh.entry_used;
        }
        if (lane == 0) {
            st_relaxed_u64(a.tile_state + tile, kStatusPrefix | exit_tag | (before + tile_total));
            sh.base = before;
            if (tile == a.num_tiles - 1) {
                *a.total = before + tile_total;
                a.entry_exit[1] = sh.last_exit;
            }
        }
    }
    __syncthreads();
    const unsigned long long out_base = sh.base;

    // ---- write: decode once more from the settled start into the staging buffer, then store
    // the staged text with aligned 16-byte writes
    uint32_t bad = 0;
    const bool fits = !TAIL && tile_total <= (uint32_t)kUnpackStageBytes && out_base + tile_total <= a.max_symbols;
    for (uint32_t chunk_lo = 0; chunk_lo < tile_total; chunk_lo += kUnpackStageBytes) {
        const unsigned long long g0 = out_base + chunk_lo;
        if (g0 >= a.max_symbols) break;
        uint32_t clen = min((uint32_t)kUnpackStageBytes, tile_total - chunk_lo);
        if (g0 + clen > a.max_symbols) clen = (uint32_t)(a.max_symbols - g0);
        uint8_t *dst = a.out + g0;
        const uint32_t align = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
        if (owned && my_cnt) {
            bool done = false;
            if (fits) {
                const uint32_t c = write_walk_fast(w, start, smem_addr(stage) + align + my_off, clut_s, wlut_s, a.nodes);
                done = !(c & kLutMarker);
            }
            if (!done && my_off < chunk_lo + clen && my_off + my_cnt > chunk_lo) {
                uint32_t cb;
                walk_generic<true>(w[0], w[1], w[2], w[3], w[4], start, lim, clut_s, wlut_s, a.nodes, &cb, stage + align,
                                   my_off - chunk_lo, clen);
                bad |= cb;
            }
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
    if ((bad & 0x80000000u) && owned) atomicOr(a.error_flags, kErrInvalidCode);
}

__global__ void __launch_bounds__(kUnpackThreads) unpack_kernel(const UnpackArgs a) {
    __shared__ __align__(16) uint32_t clut_sh[kLutSize];
    __shared__ __align__(16) uint32_t wlut_sh[kLutSize];
    __shared__ __align__(16) uint8_t stage[kUnpackStageBytes + 32];
    __shared__ TileShared sh;

    for (int i = threadIdx.x; i < kLutSize; i += kUnpackThreads) {
        clut_sh[i] = a.clut[i];
        wlut_sh[i] = a.wlut[i];
    }
    const uint32_t clut_s = smem_addr(clut_sh), wlut_s = smem_addr(wlut_sh);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            sh.tile = atomicAdd(a.ticket, 1u);
            sh.abort = ld_relaxed_u32(a.error_flags) & (kErrSeam | kErrNoConvergence);
        }
        __syncthreads();
        const uint32_t tile = sh.tile;
        if (tile >= a.num_tiles) break;
        if (sh.abort) {
            // A guess was wrong somewhere: the host will rerun the stream through the chunked
            // path.  Drain the tickets, publishing descriptors so that no look-back waits forever.
            if (threadIdx.x == 0) st_relaxed_u64(a.tile_state + tile, kStatusPrefix);
            continue;
        }
        // does the tile's last subsequence plus its 32-bit look-ahead reach past the stream end?
        const long long g_last = a.g_first + (long long)tile * kUnpackOwned + kUnpackOwned;
        if ((unsigned long long)g_last * kSubseqBits + 32 > a.end_bit)
            unpack_tile<true>(a, tile, clut_s, wlut_s, stage, sh);
        else
            unpack_tile<false>(a, tile, clut_s, wlut_s, stage, sh);
    }
}

}  // namespace

UnpackGeometry unpack_geometry(const void *d_body, size_t body_bytes) {
    UnpackGeometry g;
    const uintptr_t p = reinterpret_cast<uintptr_t>(d_body);
    const uint32_t mis = (uint32_t)(p & 15u);
    g.body_aligned = reinterpret_cast<const uint8_t *>(p - mis);
    g.byte_lo = mis;
    g.byte_hi = (uint64_t)mis + body_bytes;
    g.own_begin_bit = (uint64_t)mis * 8;
    g.own_end_bit = g.byte_hi * 8;
    g.end_bit = g.byte_hi * 8;
    g.head_known = true;
    g.head_bit = g.own_begin_bit;
    const uint64_t n_subseq = (g.own_end_bit + kSubseqBits - 1) / kSubseqBits;
    g.num_tiles = body_bytes ? (uint32_t)((n_subseq + kUnpackOwned - 1) / kUnpackOwned) : 0u;
    return g;
}

UnpackGeometry unpack_geometry_shard(const void *d_range, size_t range_bytes, size_t own_begin_byte, size_t own_end_byte,
                                     long long head_bit) {
    UnpackGeometry g;
    g.body_aligned = static_cast<const uint8_t *>(d_range);  // caller guarantees 16-byte alignment
    g.byte_lo = 0;
    g.byte_hi = range_bytes;
    g.own_begin_bit = (uint64_t)own_begin_byte * 8;
    g.own_end_bit = (uint64_t)own_end_byte * 8;
    g.end_bit = (uint64_t)range_bytes * 8;
    g.head_known = head_bit >= 0;
    g.head_bit = head_bit >= 0 ? (uint64_t)head_bit : g.own_begin_bit;
    const uint64_t first = g.own_begin_bit / kSubseqBits;
    const uint64_t end = (g.own_end_bit + kSubseqBits - 1) / kSubseqBits;
    g.num_tiles = end > first ? (uint32_t)((end - first + kUnpackOwned - 1) / kUnpackOwned) : 0u;
    return g;
}

size_t unpack_scratch_bytes(uint32_t num_tiles) { return 64 + (size_t)num_tiles * 8; }
UnpackScratch unpack_scratch_carve(void *base, uint32_t num_tiles) {
    (void)num_tiles;
    UnpackScratch s;
    uint8_t *p = static_cast<uint8_t *>(base);
    s.ticket = reinterpret_cast<uint32_t *>(p);
    s.error_flags = reinterpret_cast<uint32_t *>(p + 4);
    s.total = reinterpret_cast<unsigned long long *>(p + 8);
    s.entry_exit = reinterpret_cast<uint32_t *>(p + 24);
    s.tile_state = reinterpret_cast<unsigned long long *>(p + 64);
    return s;
}

cudaError_t launch_unpack(const UnpackGeometry &g, const uint32_t *d_clut, const uint32_t *d_wlut, const uint32_t *d_nodes,
                          uint8_t *d_out, uint64_t max_symbols, const UnpackScratch &s, void *scratch_base,
                          size_t scratch_bytes, int num_sms, cudaStream_t stream, int *launches) {
    cudaError_t err = cudaMemsetAsync(scratch_base, 0, scratch_bytes, stream);
    if (err != cudaSuccess) return err;
    if (g.num_tiles == 0) return cudaSuccess;
    UnpackArgs a;
    a.body_aligned = g.body_aligned;
    a.end_bit = g.end_bit;
    a.byte_lo = g.byte_lo;
    a.byte_hi = g.byte_hi;
    a.g_first = (long long)(g.own_begin_bit / kSubseqBits);
    a.g_own_end = (long long)((g.own_end_bit + kSubseqBits - 1) / kSubseqBits);
    a.g_max = (long long)((g.byte_hi + 15) / 16);
    a.head_known = g.head_known ? 1u : 0u;
    a.head_bit = (uint32_t)(g.head_bit - (uint64_t)a.g_first * kSubseqBits);
    a.num_tiles = g.num_tiles;
    a.clut = d_clut;
    a.wlut = d_wlut;
    a.nodes = d_nodes;
    a.out = d_out;
    a.max_symbols = max_symbols;
    a.tile_state = s.tile_state;
    a.ticket = s.ticket;
    a.error_flags = s.error_flags;
    a.total = s.total;
    a.entry_exit = s.entry_exit;
    unsigned grid = (unsigned)num_sms * 5u;  // persistent; ~45 KiB smem and 256 threads per CTA
    if (grid > g.num_tiles) grid = g.num_tiles;
    unpack_kernel<<<grid, kUnpackThreads, 0, stream>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace et
