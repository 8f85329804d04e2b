// Host half of the codec (row H1 of SURVEY §2): symbol order, two-queue tree, code
// assignment, .et header writer, dictionary parser and the decoder's lookup tables.
// Runs identically on every rank; 256 symbols, microseconds.  No CUDA in this file.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "et_internal.h"

namespace {

// MSB-first bit sink over a caller buffer — std.io.bitWriter(.big, ...) semantics
// (encode.zig:257): bits fill a byte from the top, flush pads with zeros.
class BitSink {
   public:
    BitSink(uint8_t *buf, size_t cap) : buf_(buf), cap_(cap) {}
    void put(uint64_t value, unsigned nbits) {
        for (unsigned k = nbits; k > 0; --k) bit((value >> (k - 1)) & 1u);
    }
    void bit(unsigned b) {
        if ((nbits_ >> 3) < cap_) {
            uint8_t &byte = buf_[nbits_ >> 3];
            if ((nbits_ & 7) == 0) byte = 0;
            byte |= (uint8_t)((b & 1u) << (7 - (nbits_ & 7)));
        } else {
            overflow_ = true;
        }
        ++nbits_;
    }
    void pad_to_byte() { nbits_ = (nbits_ + 7) & ~(size_t)7; }
    size_t bytes() const { return (nbits_ + 7) >> 3; }
    bool overflow() const { return overflow_; }

   private:
    uint8_t *buf_;
    size_t cap_;
    size_t nbits_ = 0;
    bool overflow_ = false;
};

// The bit the reference writes for position j (length..1) of a code (encode.zig:293,311).
inline unsigned emitted_bit(const et_code &c, unsigned j) { return (c.data >> ((j - 1) & 31u)) & 1u; }

struct BitSource {
    const uint8_t *p;
    size_t nbits;
    size_t at;
    bool take(unsigned n, uint64_t *out) {
        if (at + n > nbits) return false;
        uint64_t v = 0;
        for (unsigned k = 0; k < n; ++k, ++at) v = (v << 1) | ((p[at >> 3] >> (7 - (at & 7))) & 1u);
        *out = v;
        return true;
    }
};

}  // namespace

// ====================================================================== public host-only ABI
extern "C" int et_build_codebook(const uint64_t counts[256], et_codebook *cb) {
    if (!counts || !cb) return ET_ERR_INVALID_ARG;
    std::memset(cb, 0, sizeof *cb);

    // E2 (encode.zig:54-74): ascending count, ties by ascending byte value; zero counts
    // never enter.  The reference's write index is a u8 that saturates at 255 and is then
    // read as an exclusive length (encode.zig:70,79): with all 256 byte values present the
    // last symbol in this order gets no leaf, hence no code.
    uint8_t order[256];
    int n = 0;
    for (int s = 0; s < 256; ++s)
        if (counts[s] > 0) order[n++] = (uint8_t)s;
    std::stable_sort(order, order + n, [&](uint8_t a, uint8_t b) { return counts[a] < counts[b]; });
    if (n == 256) n = 255;
    if (n == 0) return ET_ERR_QUEUE_EMPTY;  // dequeue on an empty queue, encode.zig:138

    // E3 (encode.zig:82-138, queue.zig:9-43): two FIFO queues.  Leaves sit in slots
    // [0,n) in the order above; merged nodes are appended, so both queues are just
    // cursors into one array.  A tie between the queue heads goes to the leaf queue
    // (encode.zig:113 uses <=); the first pick becomes the left child.
    struct Node {
        uint64_t weight;
        int kid[2];
    };
    Node nodes[512];
    for (int i = 0; i < n; ++i) nodes[i] = {counts[order[i]], {-1, -1}};
    int leaf_head = 0, sapling_head = n, total = n;
    auto take = [&]() {
        const bool leaves_left = leaf_head < n, saplings_left = sapling_head < total;
        if (!saplings_left) return leaf_head++;
        if (!leaves_left) return sapling_head++;
        return nodes[leaf_head].weight <= nodes[sapling_head].weight ? leaf_head++ : sapling_head++;
    };
    while ((n - leaf_head) + (total - sapling_head) > 1) {
        const int a = take();
        const int b = take();
        nodes[total] = {nodes[a].weight + nodes[b].weight, {a, b}};
        ++total;
    }

    // E4 (encode.zig:141-214): path from the root, left appends 0, right appends 1; the
    // path register is 32 bits wide and simply loses its top bit past depth 32.
    // Children always precede their parent in `nodes`, so one reverse sweep suffices.
    uint32_t path[512];
    uint8_t depth[512];
    const int root = total - 1;
    path[root] = 0;
    depth[root] = 0;
    for (int i = root; i >= n; --i) {
        for (int side = 0; side < 2; ++side) {
            const int k = nodes[i].kid[side];
            path[k] = (path[i] << 1) | (uint32_t)side;
            depth[k] = (uint8_t)(depth[i] + 1);
        }
    }
    cb->n_symbols = (uint32_t)n;
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    for (int i = 0; i < n; ++i) {
        et_code &c = cb->code[order[i]];
        c.data = path[i];
        c.length = depth[i];
        if (c.length > 0) {
            cb->n_entries += 1;
            lo = std::min<uint32_t>(lo, c.length);
            hi = std::max<uint32_t>(hi, c.length);
            cb->body_bits += counts[order[i]] * c.length;
        }
    }
    cb->min_length = cb->n_entries ? lo : 0;
    cb->max_length = hi;
    return ET_OK;
}

extern "C" uint64_t et_shard_bits(const uint64_t counts[256], const et_codebook *cb) {
    uint64_t bits = 0;
    for (int s = 0; s < 256; ++s) bits += counts[s] * cb->code[s].length;
    return bits;
}

extern "C" size_t et_header_size(const et_codebook *cb) {
    size_t bits = 0;
    for (int s = 0; s < 256; ++s)
        if (cb->code[s].length > 0) bits += 16 + cb->code[s].length;
    return 9 + ((bits + 7) >> 3);
}

extern "C" size_t et_encode_bound(size_t n) { return 7200 + n; }  // encode.zig:253-254

extern "C" int et_write_header(const et_codebook *cb, uint64_t n, uint8_t *out, size_t cap, size_t *header_len) {
    if (!cb || !out || !header_len) return ET_ERR_INVALID_ARG;
    BitSink sink(out, cap);
    sink.put(0xe7c0de, 24);  // encode.zig:262
    sink.put(0x01, 8);       // encode.zig:266
    uint32_t entries = 0;    // encode.zig:270-275: count - 1, or 0 when there are none
    for (int s = 0; s < 256; ++s) entries += cb->code[s].length > 0;
    sink.put(entries ? entries - 1 : 0, 8);
    sink.put(n & 0xFFFFFFFFull, 32);  // encode.zig:279: low 32 bits of text.len
    for (int s = 0; s < 256; ++s) {   // encode.zig:285-297, ascending byte value
        const et_code &c = cb->code[s];
        if (c.length == 0) continue;
        sink.put((unsigned)s, 8);
        sink.put(c.length, 8);
        for (unsigned j = c.length; j > 0; --j) sink.bit(emitted_bit(c, j));
    }
    sink.pad_to_byte();  // encode.zig:298
    if (sink.overflow()) return ET_ERR_NO_SPACE;
    *header_len = sink.bytes();
    return ET_OK;
}

extern "C" int et_parse_header(const uint8_t *in, size_t n, et_dictionary *dict) {
    if (!in || !dict) return ET_ERR_INVALID_ARG;
    std::memset(dict, 0, sizeof *dict);
    if (n < 5) return ET_ERR_CORRUPT;
    dict->n_entries = (uint8_t)(in[0] + 1);  // decode.zig:34 (u8 arithmetic)
    if (dict->n_entries == 0) return ET_ERR_CORRUPT;
    dict->body_len = ((uint32_t)in[1] << 24) | ((uint32_t)in[2] << 16) | ((uint32_t)in[3] << 8) | in[4];
    BitSource src{in, n * 8, 40};
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    const uint32_t announced = dict->n_entries;
    for (uint32_t e = 0; e < announced; ++e) {  // decode.zig:66-141
        uint64_t sym = 0, len = 0, code = 0;
        // The reference's state machine simply stops when the bytes run out (decode.zig:66,135-140): the entries
        // read so far stand and the body is empty, so nothing is decoded and no error is raised.  This is also how
        // its own output for a single distinct symbol reads back (9-byte file, zero entries, encode.zig:270-275).
        bool whole = src.take(8, &sym) && src.take(8, &len);
        if (whole && src.at + len > src.nbits) whole = false;  // the code bits are not all there: the reference stops reading
        if (whole && len > 64) return ET_ERR_UNSUPPORTED;      // (the reference's [32]u8 entry is indexed out of bounds, decode.zig:124)
        whole = whole && src.take((unsigned)len, &code);
        if (!whole) {
            dict->n_entries = e;
            dict->truncated = 1;
            dict->min_length = e ? lo : 0;
            dict->max_length = hi;
            dict->body_offset = n;
            return ET_OK;
        }
        if (len == 0) return ET_ERR_CORRUPT;  // the reference indexes entry[len - 1] (decode.zig:124): out of bounds
        dict->symbol[e] = (uint8_t)sym;
        dict->length[e] = (uint8_t)len;
        dict->code[e] = code;
        lo = std::min<uint32_t>(lo, (uint32_t)len);
        hi = std::max<uint32_t>(hi, (uint32_t)len);
    }
    dict->min_length = lo;
    dict->max_length = hi;
    dict->body_offset = (src.at + 7) >> 3;  // body starts on the next byte (decode.zig:136,156)
    return ET_OK;
}

// ====================================================================== internal tables
namespace et {

int make_pack_tables(const et_codebook &cb, PackTables *t) {
    std::memset(t, 0, sizeof *t);
    t->max_length = cb.max_length;
    t->narrow_ok = cb.max_length <= kNarrowMaxLen;
    for (int s = 0; s < 256; ++s) {
        const et_code &c = cb.code[s];
        const unsigned len = c.length;
        if (len > 64) return ET_ERR_UNSUPPORTED;
        uint64_t emitted;
        if (len == 0)
            emitted = 0;
        else if (len < 32)
            emitted = c.data & ((1u << len) - 1u);
        else if (len == 32)
            emitted = c.data;
        else  // low (len-32) bits of data, then all 32 bits of data (encode.zig:311, shift is u5)
            emitted = ((uint64_t)(len == 64 ? c.data : (c.data & ((1u << (len - 32)) - 1u))) << 32) | c.data;
        t->wide_code[s] = emitted;
        t->wide_len[s] = (uint8_t)len;
        t->narrow[s][0] = t->narrow_ok ? (uint32_t)emitted : 0u;
        t->narrow[s][1] = t->narrow_ok ? len : 0u;
    }
    return ET_OK;
}

// The dictionary as a binary trie - what the device needs: it derives every lookup table from it (build_tables_kernel,
// lane_tables_kernel) - plus what validation asks about the dictionary.
// The reference decoder accepts any dictionary: it tries code lengths from the shortest up (decode.zig:175-181), so where
// one entry is a prefix of another the shorter one always matches first and the longer one can never be reached, and a
// repeated (length, code) pair overwrites the earlier symbol (decode.zig:123-125).  The same rules here: a longer code that
// runs into a leaf is dropped, a shorter code that ends on an inner node replaces the subtree, a repeat replaces the leaf;
// `prefix_free` records whether any of that happened (ET_FLAG_VALIDATE turns it into an error).
int make_unpack_trie(const et_dictionary &dict, UnpackTrie *t) {
    t->n_nodes = 0;
    t->complete = t->prefix_free = false;
    if (dict.n_entries == 0 || dict.n_entries > 256) return ET_ERR_CORRUPT;
    if (dict.max_length > 32) return ET_ERR_UNSUPPORTED;  // reference table is [32]u8 per code (decode.zig:49)
    uint16_t (*kid)[2] = t->kid;
    uint32_t n_nodes = 1;
    kid[0][0] = kid[0][1] = kChildNone;
    uint64_t kraft = 0;  // in units of 2^-32
    t->prefix_free = true;
    for (uint32_t e = 0; e < dict.n_entries; ++e) {
        const unsigned len = dict.length[e];
        if (len == 0) return ET_ERR_CORRUPT;
        kraft += 1ull << (32 - len);
        uint32_t node = 0;
        for (unsigned k = len; k > 0; --k) {
            const unsigned b = (unsigned)((dict.code[e] >> (k - 1)) & 1u);
            uint16_t &slot = kid[node][b];
            if (k == 1) {
                if (slot != kChildNone) t->prefix_free = false;
                slot = (uint16_t)(kChildLeaf | dict.symbol[e]);
            } else {
                if (slot == kChildNone) {
                    kid[n_nodes][0] = kid[n_nodes][1] = kChildNone;
                    slot = (uint16_t)n_nodes++;
                } else if (slot & kChildLeaf) {
                    t->prefix_free = false;
                    break;
                }
                node = slot;
            }
        }
    }
    t->n_nodes = n_nodes;
    t->complete = kraft == (1ull << 32);
    for (uint32_t i = 0; i < n_nodes; ++i) t->nodes[i] = ((uint32_t)kid[i][1] << 16) | kid[i][0];
    return ET_OK;
}

void format_file_size(char *buf, size_t cap, double byte_count) {
    const float b = (float)byte_count;  // utils.zig:3 takes an f32
    if (b < 1024.f)
        std::snprintf(buf, cap, "%.0f B", b);
    else if (b < 1024.f * 1024.f)
        std::snprintf(buf, cap, "%.2f KB", b / 1024.f);
    else if (b < 1024.f * 1024.f * 1024.f)
        std::snprintf(buf, cap, "%.2f MB", b / (1024.f * 1024.f));
    else
        std::snprintf(buf, cap, "%.2f GB", b / (1024.f * 1024.f * 1024.f));
}

}  // namespace et
