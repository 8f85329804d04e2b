// Chunked decoder — the path that needs no luck (decode.zig:143-203 for ANY prefix code).
//
// The single-pass kernel (et_unpack.cu) bets that a wrong start re-synchronises within a
// 128-byte warm-up.  Codes with nearly equal lengths (uniform bytes: 7- and 8-bit codes)
// need kilobytes, codes whose lengths share a factor may never re-synchronise.  Here the
// stream is cut into chunks of kChunkBytes, one THREAD per chunk, and the codeword boundary
// at which each chunk starts is found by fixpoint iteration:
//   round 0      every chunk is decoded from a guess (its first bit; chunk 0 knows the truth),
//                recording how many symbols begin in it and where its last codeword ends;
//   round r > 0  a chunk whose recorded start differs from its left neighbour's recorded end
//                decodes again from there.  A round in which nothing changed proves, by
//                induction from chunk 0, that every start is the true one.  Streams that
//                re-synchronise inside a chunk (all practical ones) need 2-4 rounds; the
//                worst case is one round per chunk, which still terminates.
//   scan         exclusive prefix sum of the symbol counts (one block, 64-bit);
//   write        every chunk decodes once more from its proven start into its final place.
// Work per symbol is one table lookup per pass with no speculation inside a chunk; the
// price is strided global access (each thread streams its own chunk through L1).
#include "et_device.cuh"
#include "et_kernels.cuh"

namespace et {

namespace {

struct ChunkArgs {
    const uint8_t *body_aligned;
    uint64_t grid_bit;            // first bit of chunk 0 (own_begin rounded down to a subsequence)
    uint64_t own_end_bit;         // symbols that begin before this bit are decoded
    uint64_t end_bit;             // no code may extend past this bit
    uint64_t byte_lo, byte_hi;    // readable bytes
    uint32_t head_off;            // start of chunk 0, bits past grid_bit (true boundary or guess)
    uint32_t n_chunks;
    const uint32_t *clut;
    const uint32_t *wlut;
    const uint32_t *nodes;
    uint16_t *start_off;  // [n] first codeword of the chunk, bits past the chunk's first bit
    uint16_t *exit_off;   // [n] first codeword boundary at or after the chunk's end, bits past that end
    uint32_t *count;      // [n] symbols that begin inside the chunk
    unsigned long long *prefix;  // [n] exclusive scan of count
    uint32_t *changed;    // [1]
    uint32_t *error_flags;
    unsigned long long *total;
    uint32_t *entry_exit;
    uint8_t *out;
    uint64_t max_symbols;
};

constexpr uint32_t kChunkBits = kChunkBytes * 8;

__device__ __forceinline__ uint32_t trie_walk(uint32_t win, uint32_t entry, const uint32_t *__restrict__ nodes,
                                              uint32_t *sym) {
    uint32_t node = entry;
    if (node == kChildNone) return 0;
    for (int b = kLutBits; b < 32; ++b) {
        const uint32_t child = (__ldg(nodes + node) >> (16 * ((win >> (31 - b)) & 1u))) & 0xFFFFu;
        if (child == kChildNone) return 0;
        if (child & kChildLeaf) {
            *sym = child & 0xFFu;
            return (uint32_t)b + 1u;
        }
        node = child;
    }
    return 0;
}

// 16 aligned stream bytes as four big-endian words; bytes outside the stream read as 0.
__device__ __forceinline__ uint4 stream_quad(const ChunkArgs &a, uint64_t qi) {
    const uint64_t byte = qi * 16;
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

// Sequential big-endian word reader over 16-byte loads.
struct WordReader {
    uint4 q;
    uint64_t qi;
    __device__ __forceinline__ void seek(const ChunkArgs &a, uint64_t wi) {
        qi = wi >> 2;
        q = stream_quad(a, qi);
    }
    __device__ __forceinline__ uint32_t word(const ChunkArgs &a, uint64_t wi) {
        if ((wi >> 2) != qi) seek(a, wi);
        const uint32_t k = (uint32_t)wi & 3u;
        return k == 0 ? q.x : k == 1 ? q.y : k == 2 ? q.z : q.w;
    }
};

// Decode from absolute bit `pos` every symbol that begins before `own_end`; nothing may end
// after `hard_end` (the end of the stream).  Returns the position reached.  WRITE stores the
// symbols at out[o..) while o < max_symbols.
template <bool WRITE>
__device__ __forceinline__ uint64_t walk_chunk(const ChunkArgs &a, uint64_t pos, uint64_t own_end, uint64_t hard_end,
                                               const uint32_t *__restrict__ clut, const uint32_t *__restrict__ wlut,
                                               uint32_t *count, uint64_t o, bool *bad) {
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
        const uint32_t c = clut[idx];
        if (!WRITE && !(c & kLutMarker) && pos + kLutBits <= own_end) {  // every code in the window begins before own_end
            pos += c & 0xffu;
            n += (c & 0xffffu) >> 9;
            continue;
        }
        uint32_t len = (c >> 16) & 0xffu, sym = wlut[idx] & 0xffu;
        if (c & kLutMarker) {
            len = trie_walk(win, wlut[idx] & 0xffffu, a.nodes, &sym);
            if (len == 0) {  // no code here (incomplete dictionary): same rule as the single-pass kernel
                *bad = true;
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

__device__ __forceinline__ void chunk_bounds(const ChunkArgs &a, uint32_t c, uint64_t *begin, uint64_t *end) {
    *begin = a.grid_bit + (uint64_t)c * kChunkBits;
    const uint64_t e = *begin + kChunkBits;
    *end = e < a.own_end_bit ? e : a.own_end_bit;
}

__global__ void __launch_bounds__(kChunkThreads) chunk_sync_kernel(const ChunkArgs a, int round) {
    __shared__ uint32_t clut_sh[kLutSize];
    const uint32_t c = blockIdx.x * kChunkThreads + threadIdx.x;
    uint32_t start = 0;
    bool work = c < a.n_chunks;
    if (work) {
        if (round == 0) {
            start = c == 0 ? a.head_off : 0u;
        } else if (c == 0) {
            work = false;
        } else {
            start = a.exit_off[c - 1];
            work = start != a.start_off[c];
        }
    }
    if (!__syncthreads_or(work)) return;  // later rounds touch only the chunks whose start moved
    for (int i = threadIdx.x; i < kLutSize; i += kChunkThreads) clut_sh[i] = a.clut[i];
    __syncthreads();
    if (!work) return;
    if (round != 0) *a.changed = 1u;
    uint64_t begin, end;
    chunk_bounds(a, c, &begin, &end);
    uint32_t cnt = 0;
    bool bad = false;
    uint64_t reached = begin + start;
    if (reached < end) reached = walk_chunk<false>(a, reached, end, a.end_bit, clut_sh, a.wlut, &cnt, 0, &bad);
    a.start_off[c] = (uint16_t)start;
    a.exit_off[c] = (uint16_t)(reached > end ? reached - end : 0);
    a.count[c] = cnt;
}

// One block: exclusive scan of count[] into prefix[], total to *total.
__global__ void __launch_bounds__(1024) chunk_scan_kernel(const ChunkArgs a) {
    __shared__ unsigned long long part[1024];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (a.n_chunks + 1023u) / 1024u;
    const uint32_t lo = min(t * per, a.n_chunks), hi = min(lo + per, a.n_chunks);
    unsigned long long sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += a.count[i];
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
        a.prefix[i] = run;
        run += a.count[i];
    }
    if (t == 1023) {
        *a.total = part[1023];
        a.entry_exit[0] = a.start_off[0];
        a.entry_exit[1] = a.exit_off[a.n_chunks - 1];
    }
}

__global__ void __launch_bounds__(kChunkThreads) chunk_write_kernel(const ChunkArgs a) {
    __shared__ uint32_t clut_sh[kLutSize];
    for (int i = threadIdx.x; i < kLutSize; i += kChunkThreads) clut_sh[i] = a.clut[i];
    __syncthreads();
    const uint32_t c = blockIdx.x * kChunkThreads + threadIdx.x;
    if (c >= a.n_chunks) return;
    const unsigned long long o = a.prefix[c];
    if (o >= a.max_symbols || a.count[c] == 0) return;
    uint64_t begin, end;
    chunk_bounds(a, c, &begin, &end);
    uint32_t cnt = 0;
    bool bad = false;
    if (begin + a.start_off[c] < end)
        walk_chunk<true>(a, begin + a.start_off[c], end, a.end_bit, clut_sh, a.wlut, &cnt, o, &bad);
    if (bad) atomicOr(a.error_flags, kErrInvalidCode);
}

}  // namespace

static uint64_t chunk_count(const UnpackGeometry &g) {
    const uint64_t grid_bit = g.own_begin_bit / kSubseqBits * kSubseqBits;
    return g.own_end_bit > grid_bit ? (g.own_end_bit - grid_bit + kChunkBits - 1) / kChunkBits : 0;
}

size_t chunked_scratch_bytes(const UnpackGeometry &g) { return 64 + (size_t)chunk_count(g) * (2 + 2 + 4 + 8) + 64; }

cudaError_t launch_unpack_chunked(const UnpackGeometry &g, const uint32_t *d_clut, const uint32_t *d_wlut,
                                  const uint32_t *d_nodes, uint8_t *d_out, uint64_t max_symbols, void *scratch_base,
                                  size_t scratch_bytes, uint32_t *h_flag, cudaStream_t stream, int *launches,
                                  uint32_t *rounds_out) {
    (void)scratch_bytes;
    const uint64_t n64 = chunk_count(g);
    uint8_t *p = static_cast<uint8_t *>(scratch_base);
    cudaError_t err = cudaMemsetAsync(p, 0, 64, stream);
    if (err != cudaSuccess) return err;
    if (n64 == 0 || g.num_tiles == 0) return cudaSuccess;
    const uint32_t n = (uint32_t)n64;
    ChunkArgs a;
    a.body_aligned = g.body_aligned;
    a.grid_bit = g.own_begin_bit / kSubseqBits * kSubseqBits;
    a.own_end_bit = g.own_end_bit;
    a.end_bit = g.end_bit;
    a.byte_lo = g.byte_lo;
    a.byte_hi = g.byte_hi;
    a.head_off = (uint32_t)(g.head_bit - a.grid_bit);
    a.n_chunks = n;
    a.clut = d_clut;
    a.wlut = d_wlut;
    a.nodes = d_nodes;
    // [ticket(4) | error flags(4) | total(8) | changed(4) | pad | entry/exit(8)]: header shared with the single-pass kernel
    a.error_flags = reinterpret_cast<uint32_t *>(p + 4);
    a.total = reinterpret_cast<unsigned long long *>(p + 8);
    a.changed = reinterpret_cast<uint32_t *>(p + 16);
    a.entry_exit = reinterpret_cast<uint32_t *>(p + 24);
    a.prefix = reinterpret_cast<unsigned long long *>(p + 64);
    a.count = reinterpret_cast<uint32_t *>(p + 64 + (size_t)n * 8);
    a.start_off = reinterpret_cast<uint16_t *>(p + 64 + (size_t)n * 12);
    a.exit_off = reinterpret_cast<uint16_t *>(p + 64 + (size_t)n * 14);
    a.out = d_out;
    a.max_symbols = max_symbols;

    const unsigned grid = (n + kChunkThreads - 1) / kChunkThreads;
    chunk_sync_kernel<<<grid, kChunkThreads, 0, stream>>>(a, 0);
    if (launches) *launches += 1;
    uint32_t rounds = 1;
    for (;;) {
        // two rounds per host check: the second finds nothing to do once the first converged
        err = cudaMemsetAsync(a.changed, 0, 4, stream);
        if (err != cudaSuccess) return err;
        chunk_sync_kernel<<<grid, kChunkThreads, 0, stream>>>(a, (int)rounds);
        chunk_sync_kernel<<<grid, kChunkThreads, 0, stream>>>(a, (int)rounds + 1);
        rounds += 2;
        if (launches) *launches += 2;
        err = cudaMemcpyAsync(h_flag, a.changed, 4, cudaMemcpyDeviceToHost, stream);
        if (err != cudaSuccess) return err;
        err = cudaStreamSynchronize(stream);
        if (err != cudaSuccess) return err;
        if (*h_flag == 0) break;
        if (rounds > n + 4u) return cudaErrorUnknown;  // cannot happen: each round settles one more chunk
    }
    chunk_scan_kernel<<<1, 1024, 0, stream>>>(a);
    chunk_write_kernel<<<grid, kChunkThreads, 0, stream>>>(a);
    if (launches) *launches += 2;
    if (rounds_out) *rounds_out = rounds;
    return cudaGetLastError();
}

}  // namespace et
