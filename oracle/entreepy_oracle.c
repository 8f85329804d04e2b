/*
 * entreepy_oracle.c — CPU restatement of typio/entreepy's Huffman encode/decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under entreepy_b200/ may link, import or call this
 * file; it is the checker used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.
 *
 * Source of truth: /root/reference/src/{encode,decode,queue}.zig (Zig 0.12, cannot be
 * compiled in this image — no Zig toolchain).  Each function cites the lines it restates.
 * PARITY UNPINNED against the reference binary: it cannot be built or run here and it ships no
 * golden .et vectors, so no byte of .et output of the real executable has been observed.
 * Pinning that does exist: the reference ships NO golden .et vectors (test.zig only checks round trips);
 * this oracle is pinned against (a) README.md:51 "477 bytes -> 374 bytes", (b) the three
 * round-trip fixtures of test.zig:35-72, (c) the hand-traced res/test.txt bytes and the
 * sha256 values in SURVEY.md §8c, which came from an independent restatement, and (d)
 * oracle/pyref.py, a second independent restatement.  Bit patterns (0/1 orientation and
 * tie order) are therefore pinned by source reading + cross-implementation agreement,
 * not by an observation of the real binary.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_ERR_QUEUE_EMPTY 1 /* queue.zig:28 via encode.zig:138 — empty input          */
#define ORACLE_ERR_NO_SPACE 2    /* fixedBufferStream NoSpaceLeft, encode.zig:254          */
#define ORACLE_ERR_CORRUPT 3
#define ORACLE_ERR_HANG 4 /* the reference decoder would spin forever */

typedef struct {
    uint32_t data;  /* encode.zig:142 — u32, silently truncates past 32 levels */
    uint8_t length; /* encode.zig:143 */
} oracle_code;

/* ---------------------------------------------------------------- E1: histogram */
/* encode.zig:43-47 */
void oracle_histogram(const uint8_t *text, size_t n, uint64_t occ[256]) {
    memset(occ, 0, 256 * sizeof(uint64_t));
    for (size_t i = 0; i < n; ++i) occ[text[i]] += 1;
}

/* ---------------------------------------------------------------- E2: symbol order */
/* encode.zig:54-74.  Ascending count, ties by ascending byte value, zero counts skipped.
 * The write index is a u8 that saturates at 255 (encode.zig:70) and is then used as an
 * EXCLUSIVE length (encode.zig:79): with 256 distinct symbols the last one in sort order
 * never becomes a leaf.  Returns symbols_length. */
int oracle_sort_symbols(const uint64_t occ[256], uint8_t sorted[256]) {
    memset(sorted, 0, 256);
    unsigned slot = 0;
    uint64_t level = 1;
    for (;;) {
        uint64_t next_level = UINT64_MAX;
        for (unsigned sym = 0; sym < 256; ++sym) {
            uint64_t c = occ[sym];
            if (c > level && c < next_level) next_level = c;
            if (c == level) {
                sorted[slot] = (uint8_t)sym;
                if (slot < 255) slot += 1;
            }
        }
        if (next_level == UINT64_MAX) break;
        level = next_level;
    }
    return (int)slot;
}

/* ---------------------------------------------------------------- E3+E4: tree and codes */
typedef struct {
    uint64_t weight;
    int left, right; /* -1 = none */
    int symbol;      /* -1 for internal nodes */
} oracle_node;

/* queue.zig:9-43 — fixed ring FIFO.  Only count/front/back semantics matter here. */
typedef struct {
    int slots[256];
    int count, front, back;
} oracle_fifo;
static void fifo_push(oracle_fifo *q, int v) {
    q->back = (q->back % 256) + 1;
    q->slots[q->back - 1] = v;
    q->count += 1;
}
static int fifo_pop(oracle_fifo *q) {
    int v = q->slots[q->front];
    q->front = (q->front + 1) % 256;
    q->count -= 1;
    return v;
}

/* encode.zig:79-214.  dict[sym] = {data,length}; symbols that never became a leaf keep
 * {0,0} (encode.zig:146). */
int oracle_build_dictionary(const uint64_t occ[256], oracle_code dict[256]) {
    uint8_t sorted[256];
    int nsym = oracle_sort_symbols(occ, sorted);
    oracle_node nodes[513];
    int nnodes = 0;
    oracle_fifo leaves, saplings;
    memset(&leaves, 0, sizeof leaves);
    memset(&saplings, 0, sizeof saplings);
    for (int i = 0; i < 256; ++i) dict[i].data = 0, dict[i].length = 0;

    for (int i = 0; i < nsym; ++i) { /* encode.zig:89-99 */
        nodes[i].symbol = sorted[i];
        nodes[i].weight = occ[sorted[i]];
        nodes[i].left = nodes[i].right = -1;
        fifo_push(&leaves, i);
    }
    nnodes = nsym;

    while (leaves.count + saplings.count > 1) { /* encode.zig:102-135 */
        int pick[2];
        for (int k = 0; k < 2; ++k) {
            if (saplings.count == 0)
                pick[k] = fifo_pop(&leaves);
            else if (leaves.count == 0)
                pick[k] = fifo_pop(&saplings);
            else if (nodes[leaves.slots[leaves.front]].weight <=
                     nodes[saplings.slots[saplings.front]].weight) /* encode.zig:113: ties -> leaf */
                pick[k] = fifo_pop(&leaves);
            else
                pick[k] = fifo_pop(&saplings);
        }
        nodes[nnodes].symbol = -1;
        nodes[nnodes].weight = nodes[pick[0]].weight + nodes[pick[1]].weight;
        nodes[nnodes].left = pick[0];  /* encode.zig:124 */
        nodes[nnodes].right = pick[1]; /* encode.zig:125 */
        fifo_push(&saplings, nnodes);
        nnodes += 1;
    }

    int root; /* encode.zig:137-138 */
    if (leaves.count > 0)
        root = fifo_pop(&leaves);
    else if (saplings.count > 0)
        root = fifo_pop(&saplings);
    else
        return ORACLE_ERR_QUEUE_EMPTY;

    /* encode.zig:161-214 — explicit stack; right child gets a 1 bit, left a 0 bit. */
    struct {
        int node;
        oracle_code path;
    } stack[513];
    int top = 0;
    stack[top].node = root;
    stack[top].path.data = 0;
    stack[top].path.length = 0;
    top = 1;
    while (top > 0) {
        int node = stack[top - 1].node;
        oracle_code path = stack[top - 1].path;
        top -= 1;
        if (nodes[node].right >= 0) {
            stack[top].node = nodes[node].right;
            stack[top].path.data = (path.data << 1) | 1u;
            stack[top].path.length = (uint8_t)(path.length + 1);
            top += 1;
        }
        if (nodes[node].left >= 0) {
            stack[top].node = nodes[node].left;
            stack[top].path.data = (path.data << 1);
            stack[top].path.length = (uint8_t)(path.length + 1);
            top += 1;
        }
        if (nodes[node].left < 0 && nodes[node].right < 0) dict[nodes[node].symbol] = path;
    }
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- bit writer */
/* std.io.bitWriter(.big, fixedBufferStream): MSB-first, flushBits zero-pads. */
typedef struct {
    uint8_t *buf;
    size_t cap, nbytes;
    uint8_t cur;
    int fill;
    int overflow;
} oracle_bitw;
static void bw_bit(oracle_bitw *w, unsigned bit) {
    w->cur = (uint8_t)((w->cur << 1) | (bit & 1u));
    if (++w->fill == 8) {
        if (w->nbytes < w->cap)
            w->buf[w->nbytes++] = w->cur;
        else
            w->overflow = 1;
        w->cur = 0;
        w->fill = 0;
    }
}
static void bw_bits(oracle_bitw *w, uint64_t v, int nbits) {
    for (int j = nbits; j > 0; --j) bw_bit(w, (unsigned)((v >> (j - 1)) & 1u));
}
static void bw_flush(oracle_bitw *w) {
    while (w->fill != 0) bw_bit(w, 0);
}
/* encode.zig:291-295 / 309-313: one writeBits(...,1) per bit, shift truncated to u5. */
static void bw_code(oracle_bitw *w, oracle_code c) {
    for (unsigned j = c.length; j > 0; --j) bw_bit(w, (c.data >> ((j - 1) & 31u)) & 1u);
}

/* ---------------------------------------------------------------- E5+E6+E7: encode */
/* encode.zig:25-337 without UI.  Writes the complete .et file (magic included) into out.
 * cap mirrors the reference scratch of 7200 + n when the caller passes that. */
int oracle_encode(const uint8_t *text, size_t n, uint8_t *out, size_t cap, size_t *out_len,
                  oracle_code dict_out[256]) {
    uint64_t occ[256];
    oracle_code dict[256];
    oracle_histogram(text, n, occ);
    int rc = oracle_build_dictionary(occ, dict);
    if (rc != ORACLE_OK) return rc;
    if (dict_out) memcpy(dict_out, dict, sizeof dict);

    oracle_bitw w = {out, cap, 0, 0, 0, 0};
    bw_bits(&w, 0xe7c0de, 24); /* encode.zig:262 */
    bw_bits(&w, 0x01, 8);      /* encode.zig:266 */
    size_t entries = 0;        /* encode.zig:270-275 */
    for (int i = 0; i < 256; ++i)
        if (dict[i].length > 0) entries += 1;
    if (entries > 0) entries -= 1;
    bw_bits(&w, entries, 8);
    bw_bits(&w, (uint64_t)n, 32); /* encode.zig:279 — low 32 bits */
    for (int i = 0; i < 256; ++i) { /* encode.zig:285-297 */
        if (dict[i].length > 0) {
            bw_bits(&w, (uint64_t)i, 8);
            bw_bits(&w, dict[i].length, 8);
            bw_code(&w, dict[i]);
        }
    }
    bw_flush(&w);                                           /* encode.zig:298 */
    for (size_t i = 0; i < n; ++i) bw_code(&w, dict[text[i]]); /* encode.zig:304-315 */
    bw_flush(&w);                                           /* encode.zig:317 */
    if (w.overflow) return ORACLE_ERR_NO_SPACE;
    *out_len = w.nbytes;
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- header parse */
typedef struct {
    uint32_t body_len;  /* decode.zig:36-42 */
    int n_entries;      /* decode.zig:34: in[0] + 1 (u8) */
    uint8_t sym[256];
    uint8_t len[256];
    uint64_t code[256]; /* usize build_bits, decode.zig:63 */
    size_t body_off;    /* offset of the body inside `in` (in = file[4..]) */
} oracle_header;

/* decode.zig:34-141: bit-serial letter(8)/len(8)/code(len) records; body starts at the
 * next byte boundary after the last record (decode.zig:136,156). `in` is file[4..]. */
int oracle_parse_header(const uint8_t *in, size_t n, oracle_header *h) {
    if (n < 5) return ORACLE_ERR_CORRUPT;
    h->n_entries = (uint8_t)(in[0] + 1);
    h->body_len = ((uint32_t)in[1] << 24) | ((uint32_t)in[2] << 16) | ((uint32_t)in[3] << 8) | in[4];
    size_t bit = 40;
    const size_t nbits = n * 8;
    int got = 0;
    while (got < h->n_entries) {
        if (bit + 16 > nbits) return ORACLE_ERR_CORRUPT;
        unsigned s = 0, l = 0;
        for (int k = 0; k < 8; ++k, ++bit) s = (s << 1) | ((in[bit >> 3] >> (7 - (bit & 7))) & 1u);
        for (int k = 0; k < 8; ++k, ++bit) l = (l << 1) | ((in[bit >> 3] >> (7 - (bit & 7))) & 1u);
        if (bit + l > nbits) return ORACLE_ERR_CORRUPT;
        uint64_t c = 0;
        for (unsigned k = 0; k < l; ++k, ++bit) c = (c << 1) | ((in[bit >> 3] >> (7 - (bit & 7))) & 1u);
        h->sym[got] = (uint8_t)s;
        h->len[got] = (uint8_t)l;
        h->code[got] = c;
        got += 1;
    }
    h->body_off = (bit + 7) >> 3;
    return ORACLE_OK;
}

/* ---------------------------------------------------------------- decode (to original) */
/* What north_star requires of decode: the ORIGINAL bytes.  Plain bit-serial trie walk over
 * the dictionary read from the header (layout per decode.zig:34-141), body_len symbols or
 * until the stream ends.  Not a restatement of decode.zig:143-203 — see oracle_decode_ref. */
int oracle_decode(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
    oracle_header h;
    int rc = oracle_parse_header(in, n, &h);
    if (rc != ORACLE_OK) return rc;
    /* trie: child[node][bit]; leaf marker in sym_of */
    int maxnodes = 1 + 256 * 64;
    int(*child)[2] = malloc(sizeof(int[2]) * (size_t)maxnodes);
    int *sym_of = malloc(sizeof(int) * (size_t)maxnodes);
    if (!child || !sym_of) return ORACLE_ERR_CORRUPT;
    int nn = 1;
    child[0][0] = child[0][1] = -1;
    sym_of[0] = -1;
    for (int e = 0; e < h.n_entries; ++e) {
        int node = 0;
        if (h.len[e] == 0 || h.len[e] > 64) { rc = ORACLE_ERR_CORRUPT; goto done; }
        for (int k = h.len[e] - 1; k >= 0; --k) {
            int b = (int)((h.code[e] >> k) & 1u);
            if (sym_of[node] >= 0) { rc = ORACLE_ERR_CORRUPT; goto done; }
            if (child[node][b] < 0) {
                child[nn][0] = child[nn][1] = -1;
                sym_of[nn] = -1;
                child[node][b] = nn++;
            }
            node = child[node][b];
        }
        if (sym_of[node] >= 0 || child[node][0] >= 0 || child[node][1] >= 0) { rc = ORACLE_ERR_CORRUPT; goto done; }
        sym_of[node] = h.sym[e];
    }
    {
        size_t produced = 0, bit = h.body_off * 8, nbits = n * 8;
        int node = 0;
        while (produced < h.body_len && bit < nbits) {
            int b = (in[bit >> 3] >> (7 - (bit & 7))) & 1;
            bit += 1;
            node = child[node][b];
            if (node < 0) { rc = ORACLE_ERR_CORRUPT; goto done; }
            if (sym_of[node] >= 0) {
                if (produced >= cap) { rc = ORACLE_ERR_NO_SPACE; goto done; }
                out[produced++] = (uint8_t)sym_of[node];
                node = 0;
            }
        }
        *out_len = produced;
    }
done:
    free(child);
    free(sym_of);
    return rc;
}

/* ---------------------------------------------------------------- decode (reference algorithm) */
/* Faithful restatement of decode.zig:13-220 INCLUDING its defects: table slot
 * value 0 means "empty" so symbol 0x00 never matches; trailing codes shorter than
 * longest_code bits are dropped; the u32 window overflows when longest_code + 7 > 32.
 * Used (a) to show reference-decoder acceptance of our .et files on inputs where that
 * decoder is well defined and (b) as the CPU baseline for decode ("port").
 * The reference uses std.AutoHashMap(usize,[32]u8) keyed by code value (decode.zig:49);
 * here an open-addressing table with the same key/payload. */
typedef struct {
    uint64_t key;
    uint8_t used;
    uint8_t by_len[32];
} oracle_slot;
#define ORACLE_MAP_CAP 1024 /* power of two, > 2*255 */
static oracle_slot *map_find(oracle_slot *map, uint64_t key, int insert) {
    uint64_t h = key * 0x9E3779B97F4A7C15ull;
    size_t i = (size_t)(h >> 54) & (ORACLE_MAP_CAP - 1);
    for (;;) {
        if (!map[i].used) {
            if (!insert) return NULL;
            map[i].used = 1;
            map[i].key = key;
            memset(map[i].by_len, 0, 32);
            return &map[i];
        }
        if (map[i].key == key) return &map[i];
        i = (i + 1) & (ORACLE_MAP_CAP - 1);
    }
}

int oracle_decode_ref(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
    oracle_header h;
    int rc = oracle_parse_header(in, n, &h);
    if (rc != ORACLE_OK) return rc;
    oracle_slot *map = calloc(ORACLE_MAP_CAP, sizeof(oracle_slot));
    if (!map) return ORACLE_ERR_CORRUPT;
    unsigned longest = 0;
    size_t shortest = SIZE_MAX;
    for (int e = 0; e < h.n_entries; ++e) { /* decode.zig:97-125 */
        unsigned l = h.len[e];
        if (l > longest) longest = l;
        if (l < shortest) shortest = l;
        if (l == 0 || l > 32) { free(map); return ORACLE_ERR_CORRUPT; } /* [32]u8 index, decode.zig:124 */
        map_find(map, h.code[e], 1)->by_len[l - 1] = h.sym[e];
    }
    uint32_t window = 0; /* decode.zig:143 */
    size_t window_len = 0, produced = 0, symbols = 0;
    for (size_t p = h.body_off; p < n; ++p) { /* decode.zig:153-159: sections are contiguous */
        window = (window << 8) | in[p];
        window_len += 8;
        while (window_len >= longest) { /* decode.zig:166 */
            size_t before = window_len, try_len = shortest;
            int stop = 0;
            while (window_len >= try_len) {
                if (symbols >= h.body_len || window_len < try_len) { stop = 1; break; }
                uint32_t mask = (uint32_t)((((uint32_t)1) << (try_len & 31)) - 1u);
                uint64_t probe = (uint64_t)((window & (mask << ((window_len - try_len) & 31))) >> ((window_len - try_len) & 63));
                oracle_slot *s = map_find(map, probe, 0);
                if (s && s->by_len[try_len - 1] > 0) {
                    if (produced >= cap) { free(map); return ORACLE_ERR_NO_SPACE; }
                    out[produced++] = s->by_len[try_len - 1]; /* decode.zig:186 */
                    symbols += 1;
                    window &= (uint32_t)((((uint32_t)1) << ((window_len - try_len) & 31)) - 1u);
                    window_len -= try_len;
                    try_len = shortest; /* decode.zig:196, then incremented below (199) */
                }
                try_len += 1;
            }
            if (stop) break;
            if (window_len == before) { /* no code matched: the reference loops forever here */
                free(map);
                *out_len = produced;
                return ORACLE_ERR_HANG;
            }
        }
    }
    free(map);
    *out_len = produced;
    return ORACLE_OK;
}

/* Sum over symbols of count*length, in bits — body size check (README.md:51 "374 bytes"). */
uint64_t oracle_body_bits(const uint64_t occ[256], const oracle_code dict[256]) {
    uint64_t bits = 0;
    for (int i = 0; i < 256; ++i) bits += occ[i] * dict[i].length;
    return bits;
}
