// K3-K5 — parallel Huffman decode of an .et body (replaces decode.zig:143-203).
//
// The stream carries no block index, so nobody knows where a codeword starts.  The body is
// cut into chunks and the codeword boundary at which each chunk starts is found by
// self-synchronisation plus a fixpoint check:
//   sync 0   every chunk starts decoding a few words BEFORE its first bit, from a guess.
//            Huffman codes re-synchronise after a few symbols, so by the time the walk
//            enters the chunk it is almost always on a true boundary.  It records where it
//            entered, how many symbols begin in the chunk and where its last codeword ends.
//   sync r   a chunk whose recorded entry differs from its left neighbour's recorded end
//            decodes again from there.  A pass in which nothing differed proves, by
//            induction from chunk 0 (true start), that every entry is the true one.  Text
//            needs one repair round; codes with nearly equal lengths (uniform bytes:
//            7/8-bit codes) need tens; the worst case is one round per chunk and still ends.
//   scan     exclusive scan of the symbol counts (64-bit);
//   write    every chunk decodes once more from its proven entry.
// Nothing here bets on luck: the guess only decides how many chunks the repair rounds redo.
//
// Two implementations of that protocol live in this file:
//   * the LANE-INTERLEAVED decoder (second half of the file) for long streams of codes that
//     re-synchronise quickly — the path the benchmarks run: chunks of 33 words, a warp per
//     region of 32 chunks staged in shared memory, flat two-lookup walks, the text of a
//     region assembled in shared memory and stored as whole 16-byte vectors;
//   * the per-THREAD chunk kernels (first half) for short streams and for codes whose lengths
//     differ by at most 2 bits: one thread per chunk of 32..4096 bytes, stream words in
//     registers 16 bytes at a time, bit position and symbol count (or output address) in ONE
//     register, table entries that are pre-packed adds for it.  Anything unusual there (ragged
//     ends of the stream, output clipped by body_len) takes the generic walker, one symbol at
//     a time with every check.
// All of them are bound by instruction issue and the integer pipe, not by HBM (ncu: profiles/),
// so the walkers are written for instruction count.
#include <cstdio>
#include <cstdlib>

#include "et_device.cuh"
#include "et_kernels.cuh"

namespace et {

namespace {

constexpr uint32_t kPosMask = 0x1ffu;  // position field of a packed walk state (bit 8 = marker)

struct DecArgs {
    const uint8_t *body_aligned;
    uint64_t grid_bit;            // first bit of chunk 0 (own_begin rounded down to a 32-byte sector)
    uint64_t own_end_bit;         // symbols that begin before this bit are decoded
    uint64_t end_bit;             // no code may extend past this bit
    uint64_t byte_lo, byte_hi;    // readable bytes
    uint32_t head_off;            // first codeword of chunk 0, bits past grid_bit (when head_known)
    uint32_t head_known;
    uint32_t n_chunks;
    uint32_t chunk_bytes;         // multiple of 16
    uint32_t fixed_len;           // > 0: every code has this length (complete code): entries in closed form, no run-up
    const uint32_t *clut;
    const uint32_t *wlut;
    const uint32_t *nodes;
    const uint16_t *slots;        // slot of every marker window (kLutSize), then the second-level tables
    uint16_t *start_off;          // [n] first codeword of the chunk, bits past the chunk's first bit
    uint16_t *exit_off;           // [n] first codeword boundary at or after the chunk's end, bits past that end
    uint32_t *count;              // [n] symbols that begin inside the chunk
    uint32_t *mid;                // [n] lane-interleaved decoder: entry of the chunk's second part | symbols of the first << 16
    unsigned long long *block_prefix;  // [ceil(n / kChunkThreads)] exclusive scan of per-block symbol counts
    uint32_t *changed;            // [1]
    uint32_t *max_sum;            // [1] symbols of the largest region
    unsigned long long *group_prefix;  // [ceil(regions / 1024)] lane-interleaved decoder: scan of the group sums
    uint32_t *work;               // [regions] lane-interleaved decoder: regions in which an entry has to be repaired
    uint32_t *work_count;         // [1]
    uint32_t *error_flags;
    unsigned long long *total;
    uint32_t *entry_exit;
    uint8_t *out;
    uint64_t max_symbols;
};

// ------------------------------------------------------------------ shared-window accessors
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Makes a value opaque to the compiler, which otherwise recomputes shared-window addresses inside the hot loops.
__device__ __forceinline__ uint32_t pinned(uint32_t x) {
    asm volatile("" : "+r"(x));
    return x;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// Loads from the decode tables: filled once per CTA before the walks, read-only afterwards, so the
// compiler may schedule these freely among the (volatile) stores of the text.
__device__ __forceinline__ uint32_t lds_tab_u32(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_tab_u16(uint32_t addr) {
    uint32_t v;
    asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ------------------------------------------------------------------ long codes
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

// Table index of the 12-bit window at the current position, as a byte offset into a u32 table.
__device__ __forceinline__ uint32_t window_offset(uint32_t hi, uint32_t lo, uint32_t c) {
    return (__funnelshift_l(lo, hi, c) >> (32 - kLutBits - 2)) & ((kLutSize - 1) << 2);
}

// The fast loops stopped on a marker: the code at the current position is longer than the
// window.  Returns the add for that one code (len | 1 << 9); bits that are no code at all
// (incomplete dictionary) are skipped one at a time and flagged.
__device__ __forceinline__ uint32_t long_code_add(uint32_t hi, uint32_t lo, uint32_t c, const uint32_t *__restrict__ wlut,
                                                  const uint32_t *__restrict__ nodes, uint32_t *sym, uint32_t *bad) {
    const uint32_t win = __funnelshift_l(lo, hi, c);
    const uint32_t len = long_code(win, __ldg(wlut + (win >> (32 - kLutBits))) & 0xffffu, nodes, sym);
    if (len) return len | (1u << 9);
    *bad = 1u;
    return 1u;
}

// ------------------------------------------------------------------ fast walkers (per-thread chunk kernels)
// Packed state c: bits 0-8 position relative to the 32-bit word being decoded (bit 8 set =
// marker entry hit), bits 9+ symbol count (count walk) or staging address (write walk).
//
// One 16-byte piece (w0..w3, w4 = first word of the next piece).  LAST: the piece ends the
// chunk, so only symbols that BEGIN before its last bit may be consumed: in the last word a
// multi-symbol window is used only while it cannot cross that bit, then single symbols.
// Returns c relative to the first word of the next piece.
// (`last` is a run-time flag so that the body exists once: eight inlined copies of these loops
// did not fit the instruction cache.)
__device__ __forceinline__ uint32_t count_piece(const uint32_t (&w)[5], uint32_t c, bool last, uint32_t clut_s,
                                                const uint32_t *__restrict__ wlut, const uint32_t *__restrict__ nodes) {
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
        const uint32_t hi = w[wi], lo = w[wi + 1];
        for (;;) {
            if (wi < 3 || !last) {
                while (!(c & 0x1e0u)) c += lds_u16(clut_s + window_offset(hi, lo, c));
            } else {
                while ((c & kPosMask) <= (uint32_t)(32 - kLutBits)) c += lds_u16(clut_s + window_offset(hi, lo, c));
                while (!(c & 0x1e0u)) c += lds_u16(clut_s + window_offset(hi, lo, c) + 2);
            }
            if (!(c & kLutMarker)) break;
            uint32_t sym, bad;
            c += long_code_add(hi, lo, c, wlut, nodes, &sym, &bad) - kLutMarker;
        }
        c -= 32u;
    }
    return c;
}

// Output side of the write walk.  Decoded symbols are shifted into a 64-bit register (newest
// byte on top).  Every fourth symbol the finished 32-bit word is stored, under a predicate (no
// branch, so the lanes of a warp stay together), into the thread's ring of 16 words in shared
// memory: word m of lane L sits at ring + ((m + 1) & 15) * 128 + L * 4, so a lane only ever
// touches its own bank and neither these stores nor the loads of the flush can conflict.
// Where the lanes of a warp meet again anyway (the end of each 32-bit stream word) whole 32-byte
// sectors leave for global memory (whole sectors: the text in flight on the GPU is larger than
// L2, a half-written sector would go to DRAM twice).
// The symbol count n in the walk state doubles as the byte index from the sector grid of the
// destination: it starts at head_skip, the bytes of the first sector that belong to the chunk
// before.  Ring capacity: after a flush fewer than 32 bytes are pending and one stream word
// yields at most 33 symbols (32 one-bit codes plus the second symbol of the last window), so
// at most 64 bytes = 16 words are ever pending.
struct OutRing {
    uint32_t ring_s;   // shared address of this lane's slot 0
    uint8_t *gsector;  // global address of the sector being assembled (32-byte aligned)
    uint32_t head_skip;
    uint32_t lo, hi;   // the last 8 symbols, newest in the top byte of hi
    uint32_t stored;   // sectors that have left for global memory
};

// shared address of word m of the ring, given (m + 1) in bits 11+ of a walk state
__device__ __forceinline__ uint32_t ring_slot_after(const OutRing &r, uint32_t c) { return ((c >> 4) & 0x780u) + r.ring_s; }
__device__ __forceinline__ uint32_t ring_slot(const OutRing &r, uint32_t m) { return (((m + 1u) & 15u) << 7) + r.ring_s; }

// Append `syms` (1 or 2 symbols in its low bytes; shift = 8 or 16; shift 0 appends nothing) and
// advance the walk state by `add` (bits | symbols << 9).
__device__ __forceinline__ uint32_t emit(uint32_t c, OutRing &r, uint32_t syms, uint32_t shift, uint32_t add) {
    r.lo = __funnelshift_r(r.lo, r.hi, shift);
    r.hi = __funnelshift_r(r.hi, syms, shift);
    const uint32_t before = c;
    c += add;
    // the symbol count crossed a multiple of 4: a word is complete (one symbol of the next word may sit on top of it)
    const uint32_t crossed = (before ^ c) & (4u << 9);
    const uint32_t w = (c & (1u << 9)) ? __funnelshift_r(r.lo, r.hi, 24) : r.hi;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.shared.u32 [%1], %2;\n\t}" ::"r"(crossed),
        "r"(ring_slot_after(r, c)), "r"(w)
        : "memory");
    return c;
}

// End of a stream word: finished sectors leave together.
__device__ __forceinline__ void flush_sectors(uint32_t c, OutRing &r) {
    while (((c >> 14) & 0x3ffffu) > r.stored) {
        const uint32_t m0 = r.stored * 8u;  // first word of the sector; 8 consecutive slots, wrapping at 16
        if (r.head_skip) {  // first sector of the chunk: its leading bytes belong to the chunk before
            for (uint32_t k = r.head_skip; k < 32u; ++k) r.gsector[k] = (uint8_t)lds_u8(ring_slot(r, m0 + (k >> 2)) + (k & 3u));
            r.head_skip = 0;
        } else {
            uint32_t v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = lds_u32(ring_slot(r, m0 + i));
            *reinterpret_cast<uint4 *>(r.gsector) = make_uint4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<uint4 *>(r.gsector + 16) = make_uint4(v[4], v[5], v[6], v[7]);
        }
        r.gsector += 32;
        r.stored += 1;
    }
}

__device__ __forceinline__ uint32_t write_piece(const uint32_t (&w)[5], uint32_t c, bool last, uint32_t wlut_s, OutRing &r,
                                                const uint32_t *__restrict__ clut, const uint32_t *__restrict__ wlut,
                                                const uint32_t *__restrict__ nodes, uint32_t *bad) {
#pragma unroll
    for (int wi = 0; wi < 4; ++wi) {
        const uint32_t hi = w[wi], lo = w[wi + 1];
        for (;;) {
            if (wi < 3 || !last) {
                while (!(c & 0x1e0u)) {
                    const uint32_t e = lds_u32(wlut_s + window_offset(hi, lo, c));
                    c = emit(c, r, e, (e >> 22) & 0x18u, e >> 16);  // a marker entry appends nothing and sets bit 8
                }
            } else {
                while ((c & kPosMask) <= (uint32_t)(32 - kLutBits)) {
                    const uint32_t e = lds_u32(wlut_s + window_offset(hi, lo, c));
                    c = emit(c, r, e, (e >> 22) & 0x18u, e >> 16);
                }
                while (!(c & 0x1e0u)) {  // one symbol at a time up to the chunk's last bit
                    const uint32_t off = window_offset(hi, lo, c);
                    const uint32_t a = __ldg(clut + (off >> 2)) >> 16;
                    if (a & kLutMarker) {
                        c |= kLutMarker;
                        break;
                    }
                    c = emit(c, r, lds_u32(wlut_s + off), 8u, a);
                }
            }
            if (!(c & kLutMarker)) break;
            uint32_t sym = 0;
            const uint32_t add = long_code_add(hi, lo, c, wlut, nodes, &sym, bad);
            c -= kLutMarker;
            c = (add != 1u) ? emit(c, r, sym, 8u, add) : c + 1u;
        }
        c -= 32u;
        flush_sectors(c, r);
    }
    return c;
}

// ------------------------------------------------------------------ stream access
__device__ __forceinline__ uint4 load_piece(const DecArgs &a, uint64_t piece) {
    // 16 aligned bytes as four big-endian words; bytes outside the readable range read as 0
    const uint64_t byte = piece * 16;
    uint4 raw;
    if (byte >= a.byte_lo && byte + 16 <= a.byte_hi) {
        raw = __ldg(reinterpret_cast<const uint4 *>(a.body_aligned + byte));
    } else {
        const long long lo = (long long)a.byte_lo - (long long)byte, hi = (long long)a.byte_hi - (long long)byte;
        raw = (hi <= 0 || lo >= 16) ? make_uint4(0, 0, 0, 0)
                                    : ld_partial_v4(a.body_aligned + byte, (int)max(lo, 0ll), (int)min(hi, 16ll));
    }
    return make_uint4(bswap32(raw.x), bswap32(raw.y), bswap32(raw.z), bswap32(raw.w));
}
__device__ __forceinline__ uint4 load_piece_fast(const DecArgs &a, uint64_t piece) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.body_aligned) + piece);
    return make_uint4(bswap32(raw.x), bswap32(raw.y), bswap32(raw.z), bswap32(raw.w));
}

// ------------------------------------------------------------------ generic walker
// Sequential big-endian word reader over guarded 16-byte loads.
struct WordReader {
    uint4 q;
    uint64_t qi;
    __device__ __forceinline__ void seek(const DecArgs &a, uint64_t wi) {
        qi = wi >> 2;
        q = load_piece(a, qi);
    }
    __device__ __forceinline__ uint32_t word(const DecArgs &a, uint64_t wi) {
        if ((wi >> 2) != qi) seek(a, wi);
        const uint32_t k = (uint32_t)wi & 3u;
        return k == 0 ? q.x : k == 1 ? q.y : k == 2 ? q.z : q.w;
    }
};

// Decode from absolute bit `pos` every symbol that begins before `own_end`; nothing may end
// after `hard_end` (the end of the stream).  Returns the position reached.  WRITE stores the
// symbols at out[o..) while o < max_symbols.
template <bool WRITE>
__device__ __noinline__ uint64_t walk_generic(const DecArgs &a, uint64_t pos, uint64_t own_end, uint64_t hard_end,
                                              uint32_t *count, uint64_t o, uint32_t *bad) {
    uint32_t n = 0;
    uint64_t wi = pos >> 5;
    WordReader rd;
    rd.seek(a, wi);
    uint32_t hi = rd.word(a, wi), lo = rd.word(a, wi + 1);
    while (pos < own_end) {
        const uint64_t need = pos >> 5;
        if (need != wi) {  // a step never consumes more than 32 bits
            wi = need;
            hi = lo;
            lo = rd.word(a, wi + 1);
        }
        const uint32_t win = __funnelshift_l(lo, hi, (uint32_t)pos & 31u);
        const uint32_t idx = win >> (32 - kLutBits);
        const uint32_t c = __ldg(a.clut + idx);
        uint32_t len = (c >> 16) & 0xffu, sym = __ldg(a.wlut + idx) & 0xffu;
        if (c & kLutMarker) {
            len = long_code(win, __ldg(a.wlut + idx) & 0xffffu, a.nodes, &sym);
            if (len == 0) {  // no code here (incomplete dictionary): skip one bit, like the fast walkers
                *bad = 1u;
                pos += 1;
                continue;
            }
        }
        if (pos + len > hard_end) break;  // final pad bits look like the start of a longer code
        if (WRITE) {
            if (o < a.max_symbols) a.out[o] = (uint8_t)sym;
            ++o;
        }
        pos += len;
        n += 1;
    }
    *count = n;
    return pos;
}

// ------------------------------------------------------------------ chunk geometry
struct Chunk {
    uint64_t begin, end;  // bits; end is clipped to own_end_bit
    bool interior;        // every piece, the piece before and the piece after are plain readable stream
};
__device__ __forceinline__ Chunk chunk_of(const DecArgs &a, uint32_t c) {
    Chunk k;
    const uint64_t bits = (uint64_t)a.chunk_bytes * 8;
    k.begin = a.grid_bit + (uint64_t)c * bits;
    const uint64_t e = k.begin + bits;
    k.end = e < a.own_end_bit ? e : a.own_end_bit;
    k.interior = e <= a.own_end_bit && e + 128 <= a.end_bit && (e >> 3) + 16 <= a.byte_hi &&
                 (k.begin >> 3) >= a.byte_lo + 16 && k.begin >= a.grid_bit + 256;
    return k;
}

// 32 aligned bytes (one DRAM sector) as two pieces of big-endian words.
struct Pair {
    uint4 a, b;
};
// The raw load and the byte swap are kept apart on purpose: the next sector is requested a
// whole sector of decoding before its first use, and nothing touches the loaded registers
// (not even the swap) until then, so the load never stalls the walk.
__device__ __forceinline__ Pair load_pair_raw(const DecArgs &a, uint64_t pair) {
    Pair p;
    p.a = __ldg(reinterpret_cast<const uint4 *>(a.body_aligned) + 2 * pair);
    p.b = __ldg(reinterpret_cast<const uint4 *>(a.body_aligned) + 2 * pair + 1);
    return p;
}
__device__ __forceinline__ uint4 swap4(uint4 v) { return make_uint4(bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w)); }
__device__ __forceinline__ Pair swap_pair(const Pair &p) {
    Pair q;
    q.a = swap4(p.a);
    q.b = swap4(p.b);
    return q;
}
__device__ __forceinline__ void prefetch_chunk_l2(const DecArgs &a, const Chunk &k) {
    const uint8_t *p = a.body_aligned + (k.begin >> 3);
    for (uint32_t off = 0; off < a.chunk_bytes + 32u; off += 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + off));
}

// Count walk of an interior chunk.  warm: start one piece early from a guess and report where
// the walk entered the chunk; else start at `start` (bits past the chunk's first bit).
// Chunks are whole 32-byte sectors; the stream is read a sector at a time, one sector ahead.
// Returns the packed state relative to the chunk's end.
__device__ __forceinline__ uint32_t count_chunk_fast(const DecArgs &a, const Chunk &k, uint32_t start, bool warm,
                                                     uint32_t clut_s, uint32_t *entry) {
    const uint64_t pair0 = k.begin >> 8;
    const uint32_t n_pairs = a.chunk_bytes >> 5;
    prefetch_chunk_l2(a, k);
    uint32_t w[5];
    uint32_t c = start;
    const uint32_t n_pieces = a.chunk_bytes >> 4;
    (void)n_pairs;
    Pair raw = load_pair_raw(a, pair0);
    if (warm) {  // only the position survives the run-up; it stops on the first boundary inside the chunk
        const uint4 pre = load_piece_fast(a, 2 * pair0 - 1);
        w[0] = pre.x; w[1] = pre.y; w[2] = pre.z; w[3] = pre.w; w[4] = bswap32(raw.a.x);
        c = count_piece(w, 0u, true, clut_s, a.wlut, a.nodes) & kPosMask;
    }
    *entry = c;
    Pair cur = raw;
#pragma unroll 1
    for (uint32_t p = 0; p < n_pieces; ++p) {
        if (!(p & 1u)) {  // a new sector: swap the one that has arrived, request the next (the piece after the chunk at the end)
            cur = swap_pair(raw);
            if (p + 2 < n_pieces)
                raw = load_pair_raw(a, pair0 + (p >> 1) + 1);
            else
                raw.a = __ldg(reinterpret_cast<const uint4 *>(a.body_aligned) + 2 * (pair0 + (p >> 1) + 1));
            w[0] = cur.a.x; w[1] = cur.a.y; w[2] = cur.a.z; w[3] = cur.a.w; w[4] = cur.b.x;
        } else {
            w[0] = cur.b.x; w[1] = cur.b.y; w[2] = cur.b.z; w[3] = cur.b.w; w[4] = bswap32(raw.a.x);
        }
        c = count_piece(w, c, p + 1 == n_pieces, clut_s, a.wlut, a.nodes);
    }
    return c;
}

__global__ void __launch_bounds__(kChunkThreads, 6) chunk_sync_kernel(const DecArgs a, int round) {
    __shared__ __align__(16) uint32_t clut_sh[kLutSize];
    const uint32_t c = blockIdx.x * kChunkThreads + threadIdx.x;
    uint32_t start = 0;
    bool work = c < a.n_chunks;
    if (work && round != 0) {
        if (c == 0) {
            work = false;
        } else {
            start = a.exit_off[c - 1];
            work = start != a.start_off[c];
        }
    }
    if (!__syncthreads_or(work)) return;  // later rounds touch only the chunks whose entry moved
    for (int i = threadIdx.x; i < kLutSize; i += kChunkThreads) clut_sh[i] = a.clut[i];
    __syncthreads();
    if (!work) return;
    if (round != 0) *a.changed = 1u;
    const Chunk k = chunk_of(a, c);
    if (c == 0 && a.head_known) start = a.head_off;
    bool known = round != 0 || (c == 0 && a.head_known);
    if (a.fixed_len && round == 0) {
        // Codes of one length never re-synchronise (a wrong phase stays wrong for ever), but they need not: every
        // boundary is a whole number of codes past the first one.  Without a known head the first entry is a guess
        // (bit 0 of the owned part); the others are consistent with it, and the caller repairs the guess.
        const uint64_t base = a.grid_bit + (a.head_known ? a.head_off : 0u);
        if (k.begin > base) start = (uint32_t)((a.fixed_len - (k.begin - base) % a.fixed_len) % a.fixed_len);
        known = true;
    }
    uint32_t cnt = 0, entry = start, exit_bits = 0;
    if (k.interior) {
        const uint32_t s = count_chunk_fast(a, k, start, !known, smem_addr(clut_sh), &entry);
        cnt = s >> 9;
        exit_bits = s & kPosMask;
    } else {
        uint64_t pos = k.begin + start;
        uint32_t bad = 0, dummy = 0;
        if (!known && k.begin >= a.byte_lo * 8 + 128) {  // same run-up as the fast path
            pos = walk_generic<false>(a, k.begin - 128, k.begin, a.end_bit, &dummy, 0, &bad);
            if (pos < k.begin) pos = k.begin;
        }
        entry = (uint32_t)(pos - k.begin);
        if (pos < k.end) pos = walk_generic<false>(a, pos, k.end, a.end_bit, &cnt, 0, &bad);
        exit_bits = pos > k.end ? (uint32_t)(pos - k.end) : 0u;
    }
    a.start_off[c] = (uint16_t)entry;
    a.exit_off[c] = (uint16_t)exit_bits;
    a.count[c] = cnt;
}

// ------------------------------------------------------------------ scan
__global__ void __launch_bounds__(kChunkThreads) chunk_sum_kernel(const DecArgs a) {
    __shared__ uint32_t warp_sum[kChunkThreads / 32];
    const uint32_t c = blockIdx.x * kChunkThreads + threadIdx.x;
    uint32_t v = c < a.n_chunks ? a.count[c] : 0u;
    // the check of the guessed entries rides along: does every chunk start where its left neighbour ended?
    if (c > 0 && c < a.n_chunks && a.start_off[c] != a.exit_off[c - 1]) *a.changed = 1u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int i = 0; i < kChunkThreads / 32; ++i) s += warp_sum[i];
        a.block_prefix[blockIdx.x] = s;
    }
}

// One block: in-place exclusive scan of `arr` (block sums, or the sums of groups of 1024 regions); total and the
// shard's entry/exit.
__global__ void __launch_bounds__(1024) chunk_scan_kernel(const DecArgs a, unsigned long long *arr, uint32_t n_blocks) {
    __shared__ unsigned long long part[1024];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (n_blocks + 1023u) / 1024u;
    const uint32_t lo = min(t * per, n_blocks), hi = min(lo + per, n_blocks);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += arr[i];
    part[t] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const unsigned long long v = t >= (uint32_t)d ? part[t - d] : 0ull;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long run = part[t] - sum;
    for (uint32_t i = lo; i < hi; ++i) {
        const unsigned long long v = arr[i];
        arr[i] = run;
        run += v;
    }
    if (t == 1023) {
        *a.total = part[1023];
        a.entry_exit[0] = a.start_off[0];
        a.entry_exit[1] = a.exit_off[a.n_chunks - 1];
    }
}

// ------------------------------------------------------------------ write
__global__ void __launch_bounds__(kChunkThreads, 5) chunk_write_kernel(const DecArgs a) {
    __shared__ __align__(16) uint32_t wlut_sh[kLutSize];
    __shared__ __align__(16) uint32_t rings[(kChunkThreads / 32) * 16 * 32];  // per warp: 16 words x 32 lanes
    __shared__ uint32_t warp_sum[kChunkThreads / 32];
    for (int i = threadIdx.x; i < kLutSize; i += kChunkThreads) wlut_sh[i] = a.wlut[i];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t c = blockIdx.x * kChunkThreads + tid;
    const bool live = c < a.n_chunks;
    const uint32_t cnt = live ? a.count[c] : 0u;
    // where this chunk's text goes: block prefix + exclusive scan inside the block
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t q = 0; q < warp; ++q) before += warp_sum[q];
    const unsigned long long o = a.block_prefix[blockIdx.x] + before + (incl - cnt);
    if (!live || cnt == 0 || o >= a.max_symbols) return;

    const Chunk k = chunk_of(a, c);
    const uint32_t start = a.start_off[c];
    uint32_t bad = 0;
    if (k.interior && o + cnt <= a.max_symbols) {
        uint8_t *dst = a.out + o;
        OutRing r;
        r.ring_s = smem_addr(rings) + warp * 2048u + lane * 4u;
        r.head_skip = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 31u);
        r.gsector = dst - r.head_skip;
        r.lo = r.hi = 0;
        r.stored = 0;
        const uint32_t head = r.head_skip;
        uint32_t s = start | (head << 9);
        const uint64_t pair0 = k.begin >> 8;
        const uint32_t n_pairs = a.chunk_bytes >> 5;
        const uint32_t wlut_s = smem_addr(wlut_sh);
        // (no L2 prefetch of the whole chunk here: with the text also streaming out, lines fetched that early are evicted before use)
        uint32_t w[5];
        const uint32_t n_pieces = a.chunk_bytes >> 4;
        (void)n_pairs;
        Pair raw = load_pair_raw(a, pair0);
        Pair cur = raw;
#pragma unroll 1
        for (uint32_t p = 0; p < n_pieces; ++p) {
            if (!(p & 1u)) {  // a new sector: swap the one that has arrived, request the next (the piece after the chunk at the end)
                cur = swap_pair(raw);
                if (p + 2 < n_pieces)
                    raw = load_pair_raw(a, pair0 + (p >> 1) + 1);
                else
                    raw.a = __ldg(reinterpret_cast<const uint4 *>(a.body_aligned) + 2 * (pair0 + (p >> 1) + 1));
                w[0] = cur.a.x; w[1] = cur.a.y; w[2] = cur.a.z; w[3] = cur.a.w; w[4] = cur.b.x;
            } else {
                w[0] = cur.b.x; w[1] = cur.b.y; w[2] = cur.b.z; w[3] = cur.b.w; w[4] = bswap32(raw.a.x);
            }
            s = write_piece(w, s, p + 1 == n_pieces, wlut_s, r, a.clut, a.wlut, a.nodes, &bad);
        }
        // the unfinished word goes to the ring, then what is left of the last sector leaves bytewise
        const uint32_t n_end = s >> 9;  // bytes from the sector grid, head included
        if (n_end & 3u) sts_u32(ring_slot(r, n_end >> 2), r.hi >> (8u * (4u - (n_end & 3u))));
        for (uint32_t kk = r.head_skip ? head : (n_end & ~31u); kk < n_end; ++kk)
            r.gsector[kk & 31u] = (uint8_t)lds_u8(ring_slot(r, kk >> 2) + (kk & 3u));
    } else {
        uint32_t n = 0;
        if (k.begin + start < k.end) walk_generic<true>(a, k.begin + start, k.end, a.end_bit, &n, o, &bad);
    }
    if (bad) atomicOr(a.error_flags, kErrInvalidCode);
}

// ====================================================================== lane-interleaved decoder
// The fast path for long streams whose code re-synchronises quickly.  Same protocol as above
// (run-up from a guess, repair rounds, fixpoint check, scan, write walk) but the unit of work
// is a REGION: 32 consecutive chunks of 33 stream words, one warp per region, one lane per chunk.
//   * the region (plus 16 bytes either side) is brought into shared memory by coalesced 16-byte
//     loads and byte-swapped once on the way in.  A chunk is 33 words, so lane L reads image word
//     33 L + j: bank (L + j) mod 32 — lanes at the same depth of their chunks never conflict;
//   * the walk is ONE flat loop: two table lookups, then at most one 32-bit refill of a 64-bit
//     window (a lookup consumes at most 12 bits, so two always fit).  The lanes of a warp meet
//     again only at the end of the chunk, not at every stream word as the per-thread walkers
//     above must (their words live in registers and cannot be indexed dynamically);
//   * the write walk assembles the text of the whole region in shared memory at its final
//     relative position (the chunks of a region are consecutive in the text as well), four
//     symbols per shared store, and the warp copies it out as aligned 16-byte vectors: no
//     partly written sector ever reaches L2 except at the two ends of a region.
constexpr uint32_t kLaneWords = 33;
constexpr uint32_t kLaneBytes = kLaneWords * 4;
constexpr uint32_t kSplitWords = 17;  // the write walk decodes words [0, 17) and [17, 33) of a chunk side by side
constexpr uint32_t kRegionBytes = 32 * kLaneBytes;            // 4224 = 33 lines of 128 bytes
constexpr uint32_t kImgBytes = 16 + kRegionBytes + 16;        // run-up of lane 0 | region | look-ahead of lane 31
constexpr uint32_t kRunupWords = 4;
constexpr int kSyncWarps = 16;
constexpr uint32_t kSubBytes = (kMaxSubTables << kSubBits) * 2;   // second-level tables in shared memory
constexpr uint32_t kTableBytes = kLutSize * 4 + kSubBytes;
constexpr uint32_t kSyncSmem = kTableBytes + kSyncWarps * kImgBytes;

struct BitBuf {
    uint32_t hi, lo, nxt;  // three consecutive stream words; the window is cut from hi:lo
    uint32_t addr;         // shared address of the word after nxt
};
__device__ __forceinline__ void buf_open(BitBuf &b, uint32_t addr0) {
    b.hi = lds_u32(addr0);
    b.lo = lds_u32(addr0 + 4);
    b.nxt = lds_u32(addr0 + 8);
    b.addr = addr0 + 12;
}
__device__ __forceinline__ void buf_shift(BitBuf &b) {
    b.hi = b.lo;
    b.lo = b.nxt;
    b.nxt = lds_u32(b.addr);
    b.addr += 4;
}
// Top 32 bits of (hi:lo) << pos, pos < 64 (one funnel shift on the 64-bit pair).
__device__ __forceinline__ uint32_t window64(const BitBuf &b, uint32_t pos) {
    return (uint32_t)(((((uint64_t)b.hi << 32) | b.lo) << (pos & 63u)) >> 32);
}
// One code longer than the first-level window at bit pos (< 32) of hi:lo.  slot_e: the marker's
// second-level slot (0x8000 | slot) or kNoSlot; sub_s: shared address of the second-level tables.
// Returns the code's length, 0 = no such code.
__device__ __forceinline__ uint32_t long_code_at(const BitBuf &b, uint32_t pos, uint32_t slot_e, uint32_t sub_s,
                                                 const DecArgs &a, uint32_t *sym) {
    const uint32_t win = __funnelshift_l(b.lo, b.hi, pos);
    if (slot_e != kNoSlot) {
        const uint32_t se = lds_tab_u16(sub_s + ((slot_e & (kMaxSubTables - 1)) << (kSubBits + 1)) +
                                    ((win >> (31 - kLutBits - kSubBits)) & ((1u << (kSubBits + 1)) - 2u)));
        if (se) {
            *sym = se & 0xffu;
            return se >> 8;
        }
    }
    return long_code(win, __ldg(a.wlut + (win >> (32 - kLutBits))) & 0xffffu, a.nodes, sym);
}
// Byte offset of the table entry (32 bit) for the window at bit pos (< 64) of hi:lo.
__device__ __forceinline__ uint32_t entry_offset(const BitBuf &b, uint32_t pos) {
    return (window64(b, pos) >> (30 - kLutBits)) & ((kLutSize - 1) << 2);
}

// Count walk over `nwords` stream words whose first word is at shared address addr0, starting at
// bit pos0 (< 32) of that word.  Consumes every symbol that begins before the end of the last
// word.  Returns the symbol count; *exit_bits = bits past that end at which the walk stopped.
// Walk state c: bits 0-6 position relative to b.hi (below 64 inside the main loop, up to 95 at the
// very end), bits 9+ symbols.  Table entries (32 bit, shared): low half = bits consumed |
// symbols << 9 for every whole code in the window, 0 = the first code is longer than the window
// (the state does not move; the second lookup of a pair then reads 0 as well); high half = the
// same for the first code alone, or for a marker 0x8000 | second-level slot / kNoSlot.
__device__ __forceinline__ uint32_t lane_count(uint32_t addr0, uint32_t nwords, uint32_t pos0, uint32_t clut_s, uint32_t sub_s,
                                               const DecArgs &a, uint32_t *exit_bits) {
    BitBuf b;
    buf_open(b, addr0);
    uint32_t c = pos0;
    const uint32_t limit = addr0 + 4u * (nwords + 1u);  // b.addr == limit: hi:lo are the last two words
    while (b.addr < limit) {
        c += lds_tab_u16(clut_s + entry_offset(b, c));
        const uint32_t e2 = lds_tab_u16(clut_s + entry_offset(b, c));
        c += e2;
        if (e2 == 0) {  // rare: a code of more than 12 bits
            if (c & 32u) {
                buf_shift(b);
                c -= 32u;
            }
            uint32_t sym;
            const uint32_t len = long_code_at(b, c & 31u, lds_tab_u16(clut_s + entry_offset(b, c) + 2u), sub_s, a, &sym);
            c += len ? (len | (1u << 9)) : 1u;
        }
        if (c & 32u) {
            buf_shift(b);
            c -= 32u;
        }
    }
    // the last words: hi:lo end 64 or 32 bits short of the chunk's end (32: a long code crossed a word at the very end)
    uint32_t end_rel = 32u * (nwords + 3u) - 8u * (b.addr - addr0);
    while ((c & 127u) + kLutBits <= end_rel) {  // whole windows that cannot cross the end
        const uint32_t e = lds_tab_u16(clut_s + entry_offset(b, c));
        if (e == 0) break;
        c += e;
    }
    for (;;) {  // one symbol at a time up to the end
        if ((c & 127u) >= end_rel) break;
        if (c & 32u) {
            buf_shift(b);
            c -= 32u;
            end_rel -= 32u;
        }
        uint32_t add = lds_tab_u16(clut_s + entry_offset(b, c) + 2u);
        if (add & 0x8000u) {
            uint32_t sym;
            const uint32_t len = long_code_at(b, c & 31u, add, sub_s, a, &sym);
            add = len ? (len | (1u << 9)) : 1u;
        }
        c += add;
    }
    *exit_bits = (c & 127u) - end_rel;
    return c >> 9;
}

// Coalesced copy of a region's stream bytes (and 16 either side) into the warp's shared image,
// big-endian words swapped to native.  Split in two so that the loads of the NEXT region can be
// in flight (in registers) while the warp walks the current one.
constexpr uint32_t kImgVecs = kImgBytes / 16;            // 266
constexpr uint32_t kImgVecsPerLane = (kImgVecs + 31) / 32;  // 9
struct RegionRegs {
    uint4 v[kImgVecsPerLane];
};
__device__ __forceinline__ void region_load(const DecArgs &a, uint64_t region_byte, uint32_t lane, RegionRegs &q) {
    const uint4 *src = reinterpret_cast<const uint4 *>(a.body_aligned + region_byte - 16);
#pragma unroll
    for (uint32_t i = 0; i < kImgVecsPerLane; ++i)
        if (i * 32 + lane < kImgVecs) q.v[i] = ld_stream_v4(src + i * 32 + lane);
}
__device__ __forceinline__ void region_store(const RegionRegs &q, uint32_t img_s, uint32_t lane) {
#pragma unroll
    for (uint32_t i = 0; i < kImgVecsPerLane; ++i)
        if (i * 32 + lane < kImgVecs) {
            const uint4 w = swap4(q.v[i]);
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(img_s + (i * 32 + lane) * 16), "r"(w.x), "r"(w.y),
                         "r"(w.z), "r"(w.w)
                         : "memory");
        }
}
__device__ __forceinline__ void region_prefetch_l2(const DecArgs &a, uint64_t region_byte, uint32_t lane) {
    const uint8_t *p = a.body_aligned + region_byte + (uint64_t)lane * 128;  // 33 lines; the last one rides on lane 0's neighbour
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    if (lane == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32 * 128));
}

struct Region {
    uint64_t begin_byte;  // first byte of the region (16-byte aligned offset from body_aligned)
    bool interior;        // the whole image is plain readable stream owned by this call, and it is not region 0
};
__device__ __forceinline__ Region region_of(const DecArgs &a, uint32_t r) {
    Region g;
    g.begin_byte = (a.grid_bit >> 3) + (uint64_t)r * kRegionBytes;
    const uint64_t end_byte = g.begin_byte + kRegionBytes;
    g.interior = r > 0 && end_byte * 8 <= a.own_end_bit && end_byte * 8 + 128 <= a.end_bit && end_byte + 16 <= a.byte_hi &&
                 g.begin_byte >= a.byte_lo + 16;
    return g;
}

// ------------------------------------------------------------------ the ends of the stream
// Regions that touch the ends of the stream (the first one, the last ones) are staged with guarded
// loads — bytes outside the readable range read as zero — and their chunks are walked from the
// same shared image.  A chunk that lies wholly inside the owned stream with at least 32 bits of
// stream after it takes the fast walkers like any other; the others (the last chunk of the
// stream, a first chunk whose known start is not in its first word) are walked one symbol at a
// time with every limit checked, from the shared tables.
__device__ __forceinline__ void region_stage_guarded(const DecArgs &a, uint64_t region_byte, uint32_t img_s, uint32_t lane) {
    for (uint32_t v = lane; v < kImgVecs; v += 32) {
        const long long byte = (long long)region_byte - 16 + 16ll * v;  // region 0 starts at byte 0: its run-up is not stream
        uint4 w = make_uint4(0, 0, 0, 0);
        if (byte >= (long long)a.byte_lo && byte + 16 <= (long long)a.byte_hi) {
            w = swap4(ld_stream_v4(a.body_aligned + byte));
        } else if (byte + 16 > (long long)a.byte_lo && byte < (long long)a.byte_hi) {
            const long long lo = (long long)a.byte_lo - byte, hi = (long long)a.byte_hi - byte;
            w = swap4(ld_partial_v4(a.body_aligned + byte, (int)(lo > 0 ? lo : 0), (int)(hi < 16 ? hi : 16)));
        }
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(img_s + v * 16), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
    }
}

struct LaneGeom {
    uint32_t own_bits;   // symbols that begin in the first own_bits bits of the chunk are this chunk's (1056 unless the owned stream ends inside)
    uint32_t hard_bits;  // no code may extend past this bit of the chunk (the end of the stream), clamped
    bool fast;           // whole chunk owned and far enough from the end of the stream for the fast walkers
    bool runup;          // the 128 bits before the chunk are stream
};
__device__ __forceinline__ LaneGeom lane_geom(const DecArgs &a, uint32_t gc) {
    LaneGeom l;
    const uint64_t begin = a.grid_bit + (uint64_t)gc * (kLaneBytes * 8);
    const uint64_t own = a.own_end_bit > begin ? a.own_end_bit - begin : 0, hard = a.end_bit > begin ? a.end_bit - begin : 0;
    l.own_bits = own < kLaneBytes * 8 ? (uint32_t)own : kLaneBytes * 8;
    l.hard_bits = hard < 4096 ? (uint32_t)hard : 4096u;
    l.fast = l.own_bits == kLaneBytes * 8 && l.hard_bits >= kLaneBytes * 8 + 32;
    l.runup = begin >= a.byte_lo * 8 + 32 * kRunupWords;
    return l;
}

// One symbol at bit `pos` of the chunk whose first word is at shared address chunk_s.  single_add: the
// first-code half of the count table's entry layout (length | 1 << 9, or a marker).  Returns the code's
// length (0 = no code here).
__device__ __forceinline__ uint32_t edge_symbol(uint32_t chunk_s, uint32_t pos, uint32_t clut_s, uint32_t sub_s, const DecArgs &a,
                                                uint32_t *sym, bool want_sym, uint32_t wlut_s) {
    BitBuf b;
    b.hi = lds_u32(chunk_s + 4u * (pos >> 5));
    b.lo = lds_u32(chunk_s + 4u * (pos >> 5) + 4u);
    b.nxt = 0;
    b.addr = 0;
    const uint32_t off = entry_offset(b, pos & 31u);
    if (want_sym) {
        const uint32_t e = lds_u32(wlut_s + off);
        if (e >= 0x10000u) {
            *sym = e & 0xffu;
            return (e >> 23) & 15u;
        }
        return long_code_at(b, pos & 31u, e, sub_s, a, sym);
    }
    const uint32_t add = lds_u16(clut_s + off + 2u);
    if (!(add & 0x8000u)) return add & 0x3fu;
    return long_code_at(b, pos & 31u, add, sub_s, a, sym);
}
// Count walk with every limit checked (same rules as walk_generic).
__device__ __noinline__ uint32_t lane_count_edge(uint32_t chunk_s, uint32_t start, const LaneGeom &l, uint32_t clut_s, uint32_t sub_s,
                                                 const DecArgs &a, uint32_t *exit_bits) {
    uint32_t pos = start, cnt = 0;
    while (pos < l.own_bits) {
        uint32_t sym;
        const uint32_t len = edge_symbol(chunk_s, pos, clut_s, sub_s, a, &sym, false, 0);
        if (len == 0) {  // no code here (incomplete dictionary): skip one bit, like every other walker
            pos += 1;
            continue;
        }
        if (pos + len > l.hard_bits) break;  // final pad bits look like the start of a longer code
        pos += len;
        cnt += 1;
    }
    *exit_bits = pos > l.own_bits ? pos - l.own_bits : 0u;
    return cnt;
}

// Which regions hold a chunk whose recorded entry is not its left neighbour's recorded exit?  One warp looks at
// four regions; the list it leaves is the work of the next repair round.
__global__ void __launch_bounds__(256) region_check_kernel(const DecArgs a, uint32_t n_regions) {
    const uint32_t lane = threadIdx.x & 31, w = (blockIdx.x * 256 + threadIdx.x) >> 5;
    bool bad[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t r = w * 4 + k, gc = r * 32 + lane;
        bad[k] = r < n_regions && gc > 0 && gc < a.n_chunks && a.exit_off[gc - 1] != a.start_off[gc];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (__any_sync(0xffffffffu, bad[k]) && lane == 0) {
            a.work[atomicAdd(a.work_count, 1u)] = w * 4 + k;
            *a.changed = 1u;
        }
}

// Persistent: grid = resident CTAs, every warp strides over the regions (round 0) or over the list
// region_check_kernel left (repair rounds); the tables are filled once per CTA.
__global__ void __launch_bounds__(kSyncWarps * 32, 2) region_sync_kernel(const DecArgs a, uint32_t n_regions, int round) {
    extern __shared__ __align__(16) uint8_t dyn[];  // count table | second-level tables | one stream image per warp
    uint32_t *clut_sh = reinterpret_cast<uint32_t *>(dyn);
    uint16_t *sub_sh = reinterpret_cast<uint16_t *>(dyn + kLutSize * 4);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t stride = gridDim.x * kSyncWarps;
    const uint32_t n_items = round == 0 ? n_regions : *a.work_count;
    if (blockIdx.x * kSyncWarps >= n_items) return;  // repair rounds: usually only a few CTAs have anything to do
    for (int i = threadIdx.x; i < kLutSize; i += kSyncWarps * 32) {
        const uint32_t e = a.clut[i];
        clut_sh[i] = (e & kLutMarker) ? ((uint32_t)(a.slots[i] == kNoSlot ? kNoSlot : (0x8000u | a.slots[i])) << 16) : e;
    }
    for (uint32_t i = threadIdx.x; i < kSubBytes / 2; i += kSyncWarps * 32) sub_sh[i] = a.slots[kLutSize + i];
    __syncthreads();
    const uint32_t clut_s = pinned(smem_addr(clut_sh)), sub_s = pinned(smem_addr(sub_sh));
    const uint32_t img_s = pinned(smem_addr(dyn) + kTableBytes + warp * kImgBytes);
    for (uint32_t item = blockIdx.x * kSyncWarps + warp; item < n_items; item += stride) {
        const uint32_t r = round == 0 ? item : a.work[item];
        const uint32_t gc = r * 32 + lane;
        uint32_t start = 0;
        bool work = gc < a.n_chunks;
        if (work && round != 0) {
            if (gc == 0) {
                work = false;
            } else {
                start = a.exit_off[gc - 1];
                work = start != a.start_off[gc];
            }
        }
        if (!__any_sync(0xffffffffu, work)) continue;
        if (gc == 0 && a.head_known) start = a.head_off;
        const bool known = round != 0 || (gc == 0 && a.head_known);
        const Region g = region_of(a, r);
        uint32_t cnt = 0, entry = start, exit_bits = 0;
        if (g.interior) {
            RegionRegs q;
            region_load(a, g.begin_byte, lane, q);
            if (round == 0 && r + stride < n_regions) region_prefetch_l2(a, g.begin_byte + (uint64_t)stride * kRegionBytes, lane);
            __syncwarp();  // the walk of the region before this one has left the image
            region_store(q, img_s, lane);
            __syncwarp();
            if (work) {
                const uint32_t chunk_s = img_s + 16 + lane * kLaneBytes;
                if (!known) {
                    uint32_t e;
                    (void)lane_count(chunk_s - 4 * kRunupWords, kRunupWords, 0u, clut_s, sub_s, a, &e);
                    entry = e;
                }
                uint32_t mid_bits;
                const uint32_t cnt_a = lane_count(chunk_s, kSplitWords, entry, clut_s, sub_s, a, &mid_bits);
                cnt = cnt_a + lane_count(chunk_s + 4 * kSplitWords, kLaneWords - kSplitWords, mid_bits, clut_s, sub_s, a, &exit_bits);
                a.mid[gc] = mid_bits | (cnt_a << 16);
            }
        } else {
            __syncwarp();
            region_stage_guarded(a, g.begin_byte, img_s, lane);
            __syncwarp();
            if (work) {
                const LaneGeom l = lane_geom(a, gc);
                const uint32_t chunk_s = img_s + 16 + lane * kLaneBytes;
                if (!known && l.runup) {
                    uint32_t e;
                    (void)lane_count(chunk_s - 4 * kRunupWords, kRunupWords, 0u, clut_s, sub_s, a, &e);
                    entry = e;
                }
                if (l.fast && entry < 32u) {
                    uint32_t mid_bits;
                    const uint32_t cnt_a = lane_count(chunk_s, kSplitWords, entry, clut_s, sub_s, a, &mid_bits);
                    cnt = cnt_a + lane_count(chunk_s + 4 * kSplitWords, kLaneWords - kSplitWords, mid_bits, clut_s, sub_s, a, &exit_bits);
                    a.mid[gc] = mid_bits | (cnt_a << 16);
                } else {
                    cnt = lane_count_edge(chunk_s, entry, l, clut_s, sub_s, a, &exit_bits);
                }
            }
        }
        if (work) {
            a.start_off[gc] = (uint16_t)entry;
            a.exit_off[gc] = (uint16_t)exit_bits;
            a.count[gc] = cnt;
        }
    }
}

// Scan of the symbols per region, three small kernels: sums of every region and of every group
// of 1024 regions (plus the largest region: it sizes the text stage of a warp); one-block scan of
// the group sums (chunk_scan_kernel); scan inside every group.
constexpr uint32_t kGroupRegions = 1024;
__global__ void __launch_bounds__(1024) region_sum_kernel(const DecArgs a, uint32_t n_regions) {
    __shared__ uint32_t warp_sum[32], warp_max[32];
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t r0 = blockIdx.x * kGroupRegions + warp * 32;
    uint32_t mine = 0;  // lane k keeps the sum of region r0 + k
    for (uint32_t k = 0; k < 32; ++k) {
        const uint32_t gc = (r0 + k) * 32 + lane;
        uint32_t v = (r0 + k < n_regions && gc < a.n_chunks) ? a.count[gc] : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == k) mine = v;
    }
    if (r0 + lane < n_regions) a.block_prefix[r0 + lane] = mine;
    uint32_t sum = mine, big = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        big = max(big, __shfl_xor_sync(0xffffffffu, big, o));
    }
    if (lane == 0) {
        warp_sum[warp] = sum;
        warp_max[warp] = big;
    }
    __syncthreads();
    if (t == 0) {
        unsigned long long s = 0;
        uint32_t m = 0;
        for (int i = 0; i < 32; ++i) {
            s += warp_sum[i];
            m = max(m, warp_max[i]);
        }
        a.group_prefix[blockIdx.x] = s;
        atomicMax(a.max_sum, m);
    }
}
__global__ void __launch_bounds__(1024) region_apply_kernel(const DecArgs a, uint32_t n_regions) {
    __shared__ uint32_t warp_sum[32];
    const uint32_t t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t r = blockIdx.x * kGroupRegions + t;
    const uint32_t v = r < n_regions ? (uint32_t)a.block_prefix[r] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += up;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += warp_sum[w];
    if (r < n_regions) a.block_prefix[r] = a.group_prefix[blockIdx.x] + before + (incl - v);
}

// ------------------------------------------------------------------ write walk of a lane
struct OutAcc {
    uint32_t lo, hi;  // the last 8 symbols, newest in the top byte of hi
    uint32_t A;       // shared address of the next text byte
};
// Append the `sh`/8 symbols in the low bytes of `syms` (sh = 0, 8 or 16).  When the text address
// crosses a multiple of 4 the finished word is stored (one symbol of the next word may already
// sit on top of it).  A lane's first word may hold bytes of the lane before it: those are
// rewritten by that lane's lane_flush() after the warp has met.
__device__ __forceinline__ void emit(OutAcc &r, uint32_t syms, uint32_t sh) {
    r.lo = __funnelshift_r(r.lo, r.hi, sh);
    r.hi = __funnelshift_r(r.hi, syms, sh);
    const uint32_t a2 = r.A + (sh >> 3);
    if ((r.A ^ a2) & 4u) sts_u32((a2 & ~3u) - 4u, __funnelshift_l(r.lo, r.hi, a2 << 3));
    r.A = a2;
}
// The same for the symbols of two table entries at once (up to four symbols, 8 x symbols in the top
// five bits of either entry): at most one word is completed.
__device__ __forceinline__ void emit_pair(OutAcc &r, uint32_t e1, uint32_t e2) {
    const uint32_t s1 = e1 >> 27, s12 = s1 + (e2 >> 27);
    const uint32_t syms = (e1 & 0xffffu) | (e2 << s1);  // bits of e2 above its symbols land above the 8 x (symbols) bits that are used
    r.lo = __funnelshift_rc(r.lo, r.hi, s12);
    r.hi = __funnelshift_rc(r.hi, syms, s12);
    const uint32_t a2 = r.A + (s12 >> 3);
    if ((r.A ^ a2) & 4u) sts_u32((a2 & ~3u) - 4u, __funnelshift_l(r.lo, r.hi, a2 << 3));
    r.A = a2;
}
// The bytes after the lane's last whole word (at most 3, never before a_begin), one at a time.
__device__ __forceinline__ void lane_flush(const OutAcc &r, uint32_t a_begin) {
    uint32_t k = r.A & 3u;
    if (r.A - a_begin < k) k = r.A - a_begin;
    for (uint32_t i = 0; i < k; ++i) {
        const uint32_t byte = (r.hi >> (8u * (4u - k + i))) & 0xffu;
        asm volatile("st.shared.u8 [%0], %1;" ::"r"(r.A - k + i), "r"(byte) : "memory");
    }
}

// Table entries of the write walk (32 bit, shared): symbols (one or two) in bits 0-15, bits consumed
// in 16-21, length of the first code in 23-26, 8 x symbols in 27-31.  A marker (first code longer
// than the window) is 0x8000 | second-level slot or kNoSlot: nothing consumed, nothing appended.
// Walk state c: bits 0-6 position relative to b.hi, the rest is noise from the adds.
struct WriteWalk {
    BitBuf b;
    uint32_t c, addr0, limit;
    OutAcc r;
};
__device__ __forceinline__ void walk_open(WriteWalk &w, uint32_t addr0, uint32_t nwords, uint32_t pos0, uint32_t text_s) {
    buf_open(w.b, addr0);
    w.c = pos0;
    w.addr0 = addr0;
    w.limit = addr0 + 4u * (nwords + 1u);  // b.addr == limit: hi:lo are the last two words
    w.r.lo = w.r.hi = 0;
    w.r.A = text_s;
}
__device__ __forceinline__ void walk_refill(WriteWalk &w) {
    if (w.c & 32u) {
        buf_shift(w.b);
        w.c -= 32u;
    }
}
// The lookup before this did not move: a code of more than 12 bits (or no code at all).
__device__ __forceinline__ void walk_long(WriteWalk &w, uint32_t wlut_s, uint32_t sub_s, const DecArgs &a, uint32_t *bad) {
    walk_refill(w);
    uint32_t sym = 0;
    const uint32_t len = long_code_at(w.b, w.c & 31u, lds_tab_u32(wlut_s + entry_offset(w.b, w.c)), sub_s, a, &sym);
    if (len) {
        w.c += len;
        emit(w.r, sym, 8u);
    } else {
        *bad = 1u;
        w.c += 1u;
    }
}
// Two lookups, then at most one refill (a lookup consumes at most 12 bits).
__device__ __forceinline__ void walk_pair(WriteWalk &w, uint32_t wlut_s, uint32_t sub_s, const DecArgs &a, uint32_t *bad) {
    const uint32_t e1 = lds_tab_u32(wlut_s + entry_offset(w.b, w.c));
    w.c += e1 >> 16;
    const uint32_t e2 = lds_tab_u32(wlut_s + entry_offset(w.b, w.c));
    w.c += e2 >> 16;
    emit_pair(w.r, e1, e2);
    if (e2 < 0x10000u) walk_long(w, wlut_s, sub_s, a, bad);
    walk_refill(w);
}
// Whatever is left of the main loop, then the last words: whole windows while they cannot cross
// the end, then one symbol at a time.
__device__ __forceinline__ void walk_finish(WriteWalk &w, uint32_t nwords, uint32_t wlut_s, uint32_t sub_s, const DecArgs &a,
                                            uint32_t *bad) {
    while (w.b.addr < w.limit) walk_pair(w, wlut_s, sub_s, a, bad);
    uint32_t end_rel = 32u * (nwords + 3u) - 8u * (w.b.addr - w.addr0);
    while ((w.c & 127u) + kLutBits <= end_rel) {
        const uint32_t e = lds_tab_u32(wlut_s + entry_offset(w.b, w.c));
        if (e < 0x10000u) break;
        w.c += e >> 16;
        emit(w.r, e, e >> 27);
    }
    for (;;) {
        if ((w.c & 127u) >= end_rel) break;
        if (w.c & 32u) {
            buf_shift(w.b);
            w.c -= 32u;
            end_rel -= 32u;
        }
        const uint32_t e = lds_tab_u32(wlut_s + entry_offset(w.b, w.c));
        uint32_t sym = e & 0xffu, len = (e >> 23) & 15u;
        if (e < 0x10000u) {
            len = long_code_at(w.b, w.c & 31u, e, sub_s, a, &sym);
            if (len == 0) {
                *bad = 1u;
                w.c += 1u;
                continue;
            }
        }
        w.c += len;
        emit(w.r, sym, 8u);
    }
}
// The write walk of one chunk: its two parts are decoded side by side (two independent dependency
// chains per lane: the walk is a chain of dependent shifts and table loads, and a warp scheduler
// with four warps cannot hide it otherwise).  chunk_s: shared address of the chunk's first word;
// text_s: shared address of its first text byte; mid = entry of part two | symbols of part one << 16.
__device__ __forceinline__ void lane_write(uint32_t chunk_s, uint32_t start, uint32_t mid, uint32_t text_s, uint32_t wlut_s,
                                           uint32_t sub_s, const DecArgs &a, uint32_t *bad, OutAcc *ra, OutAcc *rb) {
    WriteWalk wa, wb;
    const uint32_t text_b = text_s + (mid >> 16);
    walk_open(wa, chunk_s, kSplitWords, start, text_s);
    walk_open(wb, chunk_s + 4 * kSplitWords, kLaneWords - kSplitWords, mid & 0xffffu, text_b);
    while (wa.b.addr < wa.limit && wb.b.addr < wb.limit) {
        const uint32_t ea1 = lds_tab_u32(wlut_s + entry_offset(wa.b, wa.c)), eb1 = lds_tab_u32(wlut_s + entry_offset(wb.b, wb.c));
        wa.c += ea1 >> 16;
        wb.c += eb1 >> 16;
        const uint32_t ea = lds_tab_u32(wlut_s + entry_offset(wa.b, wa.c)), eb = lds_tab_u32(wlut_s + entry_offset(wb.b, wb.c));
        wa.c += ea >> 16;
        wb.c += eb >> 16;
        emit_pair(wa.r, ea1, ea);
        emit_pair(wb.r, eb1, eb);
        if ((ea < eb ? ea : eb) < 0x10000u) {  // rare: one of them met a code of more than 12 bits
            if (ea < 0x10000u) walk_long(wa, wlut_s, sub_s, a, bad);
            if (eb < 0x10000u) walk_long(wb, wlut_s, sub_s, a, bad);
        }
        walk_refill(wa);
        walk_refill(wb);
    }
    walk_finish(wa, kSplitWords, wlut_s, sub_s, a, bad);
    walk_finish(wb, kLaneWords - kSplitWords, wlut_s, sub_s, a, bad);
    *ra = wa.r;
    *rb = wb.r;
}
// Write walk with every limit checked (a chunk at the ends of the stream): one symbol at a time.
__device__ __noinline__ void lane_write_edge(uint32_t chunk_s, uint32_t start, const LaneGeom &l, uint32_t text_s, uint32_t wlut_s,
                                             uint32_t sub_s, const DecArgs &a, uint32_t *bad, OutAcc *ra) {
    OutAcc r;
    r.lo = r.hi = 0;
    r.A = text_s;
    uint32_t pos = start;
    while (pos < l.own_bits) {
        uint32_t sym = 0;
        const uint32_t len = edge_symbol(chunk_s, pos, 0, sub_s, a, &sym, true, wlut_s);
        if (len == 0) {
            *bad = 1u;
            pos += 1;
            continue;
        }
        if (pos + len > l.hard_bits) break;
        pos += len;
        emit(r, sym, 8u);
    }
    *ra = r;
}

// What the write walk of a region needs to know about it (loaded one region ahead).
struct RegionMeta {
    uint32_t cnt, start, prev_exit, mid;
    unsigned long long o_w;
    bool live;
};
__device__ __forceinline__ RegionMeta region_meta(const DecArgs &a, uint32_t r, uint32_t lane) {
    RegionMeta m;
    const uint32_t gc = r * 32 + lane;
    m.live = gc < a.n_chunks;
    m.cnt = m.live ? a.count[gc] : 0u;
    m.start = m.live ? a.start_off[gc] : 0u;
    m.mid = m.live ? a.mid[gc] : 0u;
    m.prev_exit = (m.live && gc > 0) ? a.exit_off[gc - 1] : m.start;
    m.o_w = a.block_prefix[r];
    return m;
}

// Dynamic shared memory (smem_bytes, all an SM has): write table (16 KiB) | second-level tables (8 KiB) | per warp:
// stream image (kImgBytes) + text stage (sized on the device from the largest region).
// Persistent: one CTA per SM, every warp strides over the regions; the stream bytes and the
// metadata of a warp's next region are requested before it walks the current one.
__global__ void __launch_bounds__(512, 1) region_write_kernel(const DecArgs a, uint32_t n_regions, uint32_t smem_bytes) {
    extern __shared__ __align__(16) uint8_t dyn[];
    uint32_t *wlut_sh = reinterpret_cast<uint32_t *>(dyn);
    uint16_t *sub_sh = reinterpret_cast<uint16_t *>(dyn + kLutSize * 4);
    for (uint32_t i = threadIdx.x; i < (uint32_t)kLutSize; i += blockDim.x) {
        const uint32_t e = a.wlut[i], add = e >> 16, len0 = (a.clut[i] >> 16) & 0xffu;
        wlut_sh[i] = (add & kLutMarker) ? (uint32_t)(a.slots[i] == kNoSlot ? kNoSlot : (0x8000u | a.slots[i]))
                                        : ((e & 0xffffu) | ((add & 0xffu) << 16) | (len0 << 23) | ((add >> 9) << 30));
    }
    for (uint32_t i = threadIdx.x; i < kSubBytes / 2; i += blockDim.x) sub_sh[i] = a.slots[kLutSize + i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // text stage of a warp: the largest region of this stream (found by the scan, read here so that the host
    // does not have to look at it), 15 bytes of skew, whole vectors; as many warps as then fit the SM work
    const uint32_t avail = smem_bytes - kTableBytes;
    uint32_t stage_bytes = (*a.max_sum + 15u + 15u) & ~15u;
    if (stage_bytes + kImgBytes > avail) stage_bytes = (avail - kImgBytes) & ~15u;  // regions that do not fit take the generic walker
    uint32_t warps = avail / (stage_bytes + kImgBytes);
    if (warps > (blockDim.x >> 5)) warps = blockDim.x >> 5;
    if (warp >= warps) return;
    const uint32_t stride = gridDim.x * warps;
    uint32_t r = blockIdx.x * warps + warp;
    if (r >= n_regions) return;
    const uint32_t img_s = pinned(smem_addr(dyn) + kTableBytes + warp * (kImgBytes + stage_bytes));
    const uint32_t stage_s = pinned(img_s + kImgBytes), wlut_s = pinned(smem_addr(wlut_sh)), sub_s = pinned(smem_addr(sub_sh));
    uint32_t bad = 0;

    RegionMeta m = region_meta(a, r, lane);
    Region g = region_of(a, r);
    RegionRegs q;
    if (g.interior) region_load(a, g.begin_byte, lane, q);
    for (;;) {
        const uint32_t gc = r * 32 + lane;
        const bool interior = g.interior;
        if (interior) {
            __syncwarp();  // the copy-out of the region before this one is done with the shared buffers
            region_store(q, img_s, lane);
        }
        // request the next region
        const uint32_t r_next = r + stride;
        const RegionMeta m_cur = m;
        if (r_next < n_regions) {
            m = region_meta(a, r_next, lane);
            g = region_of(a, r_next);
            if (g.interior) region_load(a, g.begin_byte, lane, q);
        }
        // does every chunk start where its left neighbour ended?  (the fixpoint check rides along)
        if (m_cur.live && gc > 0 && m_cur.prev_exit != m_cur.start) *a.changed = 1u;
        const uint32_t cnt = m_cur.cnt;
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += up;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const unsigned long long o_w = m_cur.o_w;
        if (total != 0 && o_w < a.max_symbols) {
            const unsigned long long o = o_w + (incl - cnt);
            const unsigned long long o_end = o_w + total < a.max_symbols ? o_w + total : a.max_symbols;
            uint8_t *dst_w = a.out + o_w;
            const uint32_t skew = (uint32_t)(reinterpret_cast<uintptr_t>(dst_w) & 15u);  // text of other regions in the first vector
            if (skew + total <= stage_bytes) {
                if (!interior) {
                    __syncwarp();
                    region_stage_guarded(a, (a.grid_bit >> 3) + (uint64_t)r * kRegionBytes, img_s, lane);
                }
                __syncwarp();
                const uint32_t text_s = stage_s + skew + (incl - cnt), chunk_s = img_s + 16 + lane * kLaneBytes;
                OutAcc ra, rb;
                ra.lo = ra.hi = rb.lo = rb.hi = 0;
                ra.A = rb.A = text_s;
                uint32_t text_b = text_s;  // where the second accumulator started
                if (interior) {
                    text_b = text_s + (m_cur.mid >> 16);
                    lane_write(chunk_s, m_cur.start, m_cur.mid, text_s, wlut_s, sub_s, a, &bad, &ra, &rb);
                } else if (m_cur.live && cnt) {
                    const LaneGeom l = lane_geom(a, gc);
                    if (l.fast && m_cur.start < 32u) {
                        text_b = text_s + (m_cur.mid >> 16);
                        lane_write(chunk_s, m_cur.start, m_cur.mid, text_s, wlut_s, sub_s, a, &bad, &ra, &rb);
                    } else {
                        lane_write_edge(chunk_s, m_cur.start, l, text_s, wlut_s, sub_s, a, &bad, &ra);
                        rb.A = text_b = ra.A;
                    }
                }
                // the bytes after the last whole word of either part leave once every lane has stored its whole words
                __syncwarp();
                lane_flush(ra, text_s);
                lane_flush(rb, text_b);
                __syncwarp();
                // the stage is an image of the text from the 16-byte boundary below dst_w: whole vectors leave as such
                uint8_t *base = dst_w - skew;
                const uint32_t first = skew, last = skew + (uint32_t)(o_end - o_w);
                const uint32_t n_vec = (last + 15u) >> 4;
                for (uint32_t v = lane; v < n_vec; v += 32) {
                    const uint32_t b0 = v * 16u;
                    if (b0 >= first && b0 + 16u <= last) {
                        uint4 w;
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                                     : "r"(stage_s + b0));
                        st_stream_v4(base + b0, w);
                    } else {
                        const uint32_t lo = b0 > first ? b0 : first, hi = b0 + 16u < last ? b0 + 16u : last;
                        for (uint32_t k = lo; k < hi; ++k) base[k] = (uint8_t)lds_u8(stage_s + k);
                    }
                }
            } else if (m_cur.live && cnt) {  // a region whose text does not fit the stage (more than 227 KiB of shared memory holds)
                const Chunk k = chunk_of(a, gc);
                uint32_t n = 0;
                if (k.begin + m_cur.start < k.end) walk_generic<true>(a, k.begin + m_cur.start, k.end, a.end_bit, &n, o, &bad);
            }
        }
        if (r_next >= n_regions) break;
        r = r_next;
    }
    if (bad) atomicOr(a.error_flags, kErrInvalidCode);
}

uint64_t chunk_count(const UnpackGeometry &g, uint32_t chunk_bytes) {
    const uint64_t grid_bit = g.own_begin_bit / 256 * 256;
    const uint64_t bits = (uint64_t)chunk_bytes * 8;
    return g.own_end_bit > grid_bit ? (g.own_end_bit - grid_bit + bits - 1) / bits : 0;
}

}  // namespace

// Once per context (the shared-memory opt-in of the two big kernels is a property of the function on the
// context's device): what the device offers, and the attributes the launches rely on.
cudaError_t unpack_init_device(int device, UnpackTuning *tune) {
    cudaError_t err;
    if ((err = cudaDeviceGetAttribute(&tune->num_sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) return err;
    if ((err = cudaDeviceGetAttribute(&tune->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess) return err;
    if ((err = cudaFuncSetAttribute(region_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tune->max_smem)) != cudaSuccess)
        return err;
    return cudaFuncSetAttribute(region_sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyncSmem);
}

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
    return g;
}

// Streams of at least this many bytes take the lane-interleaved decoder (enough regions to give every
// SM a full set of warps).  ET_TUNE_LANE_MIN_BYTES overrides it (tests run the path on small inputs).
static uint64_t lane_path_min_bytes(const UnpackTuning &tune) {
    if (tune.lane_min_bytes >= 0)
        return (uint64_t)tune.lane_min_bytes > 2 * kRegionBytes ? (uint64_t)tune.lane_min_bytes : 2 * kRegionBytes;
    return (uint64_t)tune.num_sms * 8 * kRegionBytes;
}

// Chunk size.  256 B suits codes that re-synchronise within a few symbols.  When all code
// lengths are (nearly) equal a wrong parse survives for kilobytes (uniform bytes: 7/8-bit
// codes, ~2 KB on average), and every fixpoint round repairs only one chunk's worth of it, so
// such codes get chunks longer than their synchronisation distance.  Short streams get
// shorter chunks so that the GPU is not left mostly idle.
uint32_t unpack_chunk_bytes(const UnpackGeometry &g, const UnpackTuning &tune, uint32_t min_length, uint32_t max_length) {
    const int num_sms = tune.num_sms;
    const uint64_t bytes = (g.own_end_bit - g.own_begin_bit + 7) / 8;
    const uint32_t spread = max_length - min_length;
    if (spread > 2 && bytes >= lane_path_min_bytes(tune)) return kLaneBytes;  // lane-interleaved decoder
    uint32_t cb = spread <= 1 ? 4096u : spread == 2 ? 1024u : 256u;
    const uint32_t floor_cb = cb > 256u ? 256u : 32u;
    while (cb > floor_cb && bytes / cb < (uint64_t)num_sms * 2048) cb >>= 1;
    return cb;
}

size_t unpack_scratch_bytes(const UnpackGeometry &g, uint32_t chunk_bytes) {
    const uint64_t n = chunk_count(g, chunk_bytes);
    const uint64_t per = chunk_bytes == kLaneBytes ? 32 : kChunkThreads;  // chunks per scanned sum
    const uint64_t nb = (n + per - 1) / per;
    const uint64_t ng = (nb + kGroupRegions - 1) / kGroupRegions;
    return 64 + (size_t)nb * 8 + (size_t)n * (4 + 2 + 2 + 4) + 64 + (size_t)ng * 8 + 64 + (size_t)nb * 4 + 64;
}

// Lane-interleaved decoder: the same protocol as below with regions of 32 chunks per warp.  One extra
// look at the scratch header after the scan: the largest region sizes the text stage of a warp.
static cudaError_t launch_unpack_lanes(const DecArgs &a, uint32_t n_regions, const void *d_header, uint8_t *h_hdr, cudaStream_t stream,
                                       const UnpackTuning &tune, int *launches, uint32_t *rounds_out) {
    cudaError_t err;
    const int max_smem = tune.max_smem, num_sms = tune.num_sms;
    const uint32_t *h_changed = reinterpret_cast<const uint32_t *>(h_hdr + 16);
    const uint32_t sync_blocks = (n_regions + kSyncWarps - 1) / kSyncWarps;
    const uint32_t resident = (uint32_t)num_sms * 2u;  // __launch_bounds__(.., 2)
    const uint32_t sync_grid = sync_blocks < resident ? sync_blocks : resident;
    const uint32_t check_grid = (n_regions + 31u) / 32u;  // 8 warps x 4 regions per CTA
    // one repair round: list the regions with a wrong entry, walk those again from their neighbours' exits
    auto repair = [&](int round) -> cudaError_t {
        cudaError_t e = cudaMemsetAsync(a.work_count, 0, 4, stream);
        if (e != cudaSuccess) return e;
        region_check_kernel<<<check_grid, 256, 0, stream>>>(a, n_regions);
        region_sync_kernel<<<sync_grid, kSyncWarps * 32, kSyncSmem, stream>>>(a, n_regions, round);
        if (launches) *launches += 2;
        return cudaSuccess;
    };
    region_sync_kernel<<<sync_grid, kSyncWarps * 32, kSyncSmem, stream>>>(a, n_regions, 0);
    if (launches) *launches += 1;
    if ((err = repair(1)) != cudaSuccess) return err;
    if ((err = cudaMemsetAsync(a.changed, 0, 8, stream)) != cudaSuccess) return err;  // changed and max_sum
    uint32_t rounds = 2;
    for (;;) {
        const uint32_t n_groups = (n_regions + kGroupRegions - 1) / kGroupRegions;
        region_sum_kernel<<<n_groups, 1024, 0, stream>>>(a, n_regions);
        chunk_scan_kernel<<<1, 1024, 0, stream>>>(a, a.group_prefix, n_groups);
        region_apply_kernel<<<n_groups, 1024, 0, stream>>>(a, n_regions);
        const uint32_t write_blocks = (n_regions + 15u) / 16u;
        region_write_kernel<<<write_blocks < (uint32_t)num_sms ? write_blocks : (uint32_t)num_sms, 512, max_smem, stream>>>(
            a, n_regions, (uint32_t)max_smem);
        if (launches) *launches += 4;
        // one look at the scratch header: error flags, symbols found, "an entry was wrong", entry and exit of the shard
        if ((err = cudaMemcpyAsync(h_hdr, d_header, 32, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return err;
        if ((err = cudaStreamSynchronize(stream)) != cudaSuccess) return err;
        if (tune.debug)
            fprintf(stderr, "[lanes] regions=%u chunks=%u max_sum=%u rounds=%u changed=%u\n", n_regions, a.n_chunks,
                    *reinterpret_cast<const uint32_t *>(h_hdr + 20), rounds, *h_changed);
        if (*h_changed == 0) break;  // every entry was the true one: what the write walk produced stands
        // entries still moving: fixpoint rounds, four per host visit; a check that lists nothing is the proof
        for (;;) {
            for (int i = 0; i < 4; ++i)
                if ((err = repair((int)rounds + i)) != cudaSuccess) return err;
            rounds += 4;
            if ((err = cudaMemsetAsync(a.changed, 0, 8, stream)) != cudaSuccess) return err;
            if ((err = cudaMemsetAsync(a.work_count, 0, 4, stream)) != cudaSuccess) return err;
            region_check_kernel<<<check_grid, 256, 0, stream>>>(a, n_regions);
            if (launches) *launches += 1;
            if ((err = cudaMemcpyAsync(h_hdr, d_header, 32, cudaMemcpyDeviceToHost, stream)) != cudaSuccess) return err;
            if ((err = cudaStreamSynchronize(stream)) != cudaSuccess) return err;
            if (tune.debug) fprintf(stderr, "[lanes] repair rounds=%u changed=%u\n", rounds, *h_changed);
            if (*h_changed == 0) break;
            if (rounds > a.n_chunks + 8u) return cudaErrorUnknown;  // cannot happen: each round settles one more chunk
        }
        if ((err = cudaMemsetAsync(a.error_flags, 0, 4, stream)) != cudaSuccess) return err;  // raised by a wrong parse
    }
    if (rounds_out) *rounds_out = rounds;
    return cudaGetLastError();
}

cudaError_t launch_unpack(const UnpackGeometry &g, uint32_t chunk_bytes, const uint32_t *d_clut, const uint32_t *d_wlut,
                          const uint32_t *d_nodes, const uint16_t *d_slots, uint8_t *d_out, uint64_t max_symbols, void *scratch_base,
                          uint8_t *h_hdr, cudaStream_t stream, const UnpackTuning &tune, uint32_t fixed_len, int *launches,
                          uint32_t *rounds_out) {
    uint32_t *h_flag = reinterpret_cast<uint32_t *>(h_hdr + 32);  // a word for the check rounds
    const uint64_t n64 = chunk_count(g, chunk_bytes);
    uint8_t *p = static_cast<uint8_t *>(scratch_base);
    cudaError_t err = cudaMemsetAsync(p, 0, 64, stream);
    if (err != cudaSuccess) return err;
    if (rounds_out) *rounds_out = 0;
    if (n64 == 0) {  // nothing to decode: an all-zero header
        for (int i = 0; i < 32; ++i) h_hdr[i] = 0;
        return cudaStreamSynchronize(stream);
    }
    const uint32_t n = (uint32_t)n64;
    const bool lanes = chunk_bytes == kLaneBytes;
    const uint32_t nb = lanes ? (n + 31u) / 32u : (n + kChunkThreads - 1) / kChunkThreads;
    DecArgs a;
    a.body_aligned = g.body_aligned;
    a.grid_bit = g.own_begin_bit / 256 * 256;  // chunks are whole 32-byte sectors
    a.own_end_bit = g.own_end_bit;
    a.end_bit = g.end_bit;
    a.byte_lo = g.byte_lo;
    a.byte_hi = g.byte_hi;
    a.head_known = g.head_known ? 1u : 0u;
    a.head_off = (uint32_t)(g.head_bit - a.grid_bit);
    a.n_chunks = n;
    a.chunk_bytes = chunk_bytes;
    a.fixed_len = fixed_len;
    a.clut = d_clut;
    a.wlut = d_wlut;
    a.nodes = d_nodes;
    a.slots = d_slots;
    // [pad(4) | error flags(4) | total(8) | changed(4) | pad(4) | entry/exit(8)] then the arrays
    a.error_flags = reinterpret_cast<uint32_t *>(p + 4);
    a.total = reinterpret_cast<unsigned long long *>(p + 8);
    a.changed = reinterpret_cast<uint32_t *>(p + 16);
    a.max_sum = reinterpret_cast<uint32_t *>(p + 20);
    a.entry_exit = reinterpret_cast<uint32_t *>(p + 24);
    a.block_prefix = reinterpret_cast<unsigned long long *>(p + 64);
    a.count = reinterpret_cast<uint32_t *>(p + 64 + (size_t)nb * 8);
    a.start_off = reinterpret_cast<uint16_t *>(p + 64 + (size_t)nb * 8 + (size_t)n * 4);
    a.exit_off = reinterpret_cast<uint16_t *>(p + 64 + (size_t)nb * 8 + (size_t)n * 6);
    a.mid = reinterpret_cast<uint32_t *>(p + 64 + (size_t)nb * 8 + (size_t)n * 8);
    a.group_prefix = reinterpret_cast<unsigned long long *>(p + ((64 + (size_t)nb * 8 + (size_t)n * 12 + 63) & ~(size_t)63));
    a.work = reinterpret_cast<uint32_t *>(a.group_prefix + (nb + 1023) / 1024 + 1);
    a.work_count = reinterpret_cast<uint32_t *>(p + 32);
    a.out = d_out;
    a.max_symbols = max_symbols;

    if (lanes) return launch_unpack_lanes(a, nb, p, h_hdr, stream, tune, launches, rounds_out);

    // Common case (codes that re-synchronise quickly): one walk from the guesses, one repair
    // round for the few chunks whose run-up was too short (text: 0.1 % of them; their exits do
    // not move, because a walk that missed 128 bits of run-up still locks on inside 2048 bits of
    // chunk), the final check fused into the block sums, the write walk — and a single look at
    // the flag at the very end.  Only when that check failed (slowly synchronising codes) do
    // the fixpoint rounds run, and the sums and the write walk are repeated.
    chunk_sync_kernel<<<nb, kChunkThreads, 0, stream>>>(a, 0);
    chunk_sync_kernel<<<nb, kChunkThreads, 0, stream>>>(a, 1);
    err = cudaMemsetAsync(a.changed, 0, 4, stream);
    if (err != cudaSuccess) return err;
    if (launches) *launches += 2;
    uint32_t rounds = 2;
    // Fixpoint rounds, four per host visit (the flag is cleared before the last of them: a round
    // that changed nothing is the proof).
    auto settle = [&]() -> cudaError_t {
        for (;;) {
            for (int i = 0; i < 3; ++i) chunk_sync_kernel<<<nb, kChunkThreads, 0, stream>>>(a, (int)rounds + i);
            cudaError_t e = cudaMemsetAsync(a.changed, 0, 4, stream);
            if (e != cudaSuccess) return e;
            chunk_sync_kernel<<<nb, kChunkThreads, 0, stream>>>(a, (int)rounds + 3);
            rounds += 4;
            if (launches) *launches += 4;
            e = cudaMemcpyAsync(h_flag, a.changed, 4, cudaMemcpyDeviceToHost, stream);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(stream);
            if (e != cudaSuccess) return e;
            if (*h_flag == 0) return cudaSuccess;
            if (rounds > n + 8u) return cudaErrorUnknown;  // cannot happen: each round settles one more chunk
        }
    };
    // Long chunks were chosen because this code synchronises slowly: the guesses are known to be
    // poor, so settle the entries before spending a write walk on them.
    if (chunk_bytes > 256u && !fixed_len) {
        err = settle();
        if (err != cudaSuccess) return err;
    }
    for (;;) {
        chunk_sum_kernel<<<nb, kChunkThreads, 0, stream>>>(a);
        chunk_scan_kernel<<<1, 1024, 0, stream>>>(a, a.block_prefix, nb);
        chunk_write_kernel<<<nb, kChunkThreads, 0, stream>>>(a);
        if (launches) *launches += 3;
        err = cudaMemcpyAsync(h_hdr, p, 32, cudaMemcpyDeviceToHost, stream);  // the whole scratch header, for the caller as well
        if (err != cudaSuccess) return err;
        err = cudaStreamSynchronize(stream);
        if (err != cudaSuccess) return err;
        if (*reinterpret_cast<const uint32_t *>(h_hdr + 16) == 0) break;  // every entry was the true one: what the write walk produced stands
        err = settle();
        if (err != cudaSuccess) return err;
        // the error flags the first write walk may have raised came from a wrong parse
        err = cudaMemsetAsync(a.error_flags, 0, 4, stream);
        if (err != cudaSuccess) return err;
    }
    if (rounds_out) *rounds_out = rounds;
    return cudaGetLastError();
}

}  // namespace et
