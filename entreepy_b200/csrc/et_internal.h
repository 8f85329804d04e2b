// Internal declarations shared by the host codec (et_host.cpp), the kernels (et_*.cu) and
// the C-ABI layer (et_api.cu).  Nothing here is part of the public ABI.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/entreepy_b200.h"

namespace et {

// ---------------------------------------------------------------- encoder tables (host -> device)
// Bits the reference actually emits for Code{data,length}: for j = length..1 it writes
// (data >> ((j-1) mod 32)) & 1  (encode.zig:293,311).  For length <= 32 that is the code
// itself; for 32 < length <= 64 it is the low (length-32) bits of data followed by all 32
// bits of data (the reference's truncation artefact) — reproduced bit for bit.
struct PackTables {
    uint32_t narrow[256][2];  // {emitted bits, length}, valid when max_length <= kNarrowMaxLen
    uint64_t wide_code[256];
    uint8_t wide_len[256];
    uint32_t max_length;
    bool narrow_ok;
};
constexpr uint32_t kNarrowMaxLen = 32;

int make_pack_tables(const et_codebook &cb, PackTables *t);

// ---------------------------------------------------------------- decoder tables (built on the device from the trie)
// Two first-level tables indexed by a kLutBits-bit window of the stream.  Entries are
// pre-packed "adds" for a walk state that keeps the bit position in bits 0-8 and a symbol
// count (or output address) in bits 9+:   add = bits_consumed | symbols << 9.
// kLutMarker (bit 8) instead of an add means "the first code here is longer than the window,
// or no code at all": the fast loops stop and the generic walker takes over.
//   clut[w]  low 16: every whole code that fits in the window (1..12 symbols)
//            high 16: the first code only
//   wlut[w]  low 16: sym0 | sym1 << 8          (marker: trie node reached, 0xFFFF = no such code)
//            high 16: the first two codes if both fit, else the first
// Second level: binary trie, node = (child1 << 16) | child0; child < 0x8000 = node index,
// 0x8000|sym = leaf, 0xFFFF = no such code.
constexpr int kLutBits = 12;
constexpr int kLutSize = 1 << kLutBits;
constexpr uint32_t kLutMarker = 0x100u;
constexpr uint32_t kMaxTrieNodes = 8192;
constexpr uint32_t kChildLeaf = 0x8000u;
constexpr uint32_t kChildNone = 0xFFFFu;

// Third table (lane-interleaved decoder): codes of kLutBits+1 .. kLutBits+8 bits without a trie walk.
// Every first-level window that is a proper prefix of longer codes ("marker" entry) gets a slot;
// sub[slot][next 8 bits] = symbol | length << 8, 0 = longer than that or no code.  Dictionaries with
// more than kMaxSubTables such windows keep the trie for the rest.
constexpr uint32_t kSubBits = 8;
constexpr uint32_t kMaxSubTables = 16;
constexpr uint16_t kNoSlot = 0xFFFFu;

struct UnpackTrie {
    uint32_t nodes[kMaxTrieNodes];
    uint16_t kid[kMaxTrieNodes][2];  // scratch of the builder
    uint32_t n_nodes;
    bool complete, prefix_free;
};
int make_unpack_trie(const et_dictionary &dict, UnpackTrie *t);

// "{d} B" / "{d:.2} KB" ... of utils.zig:3-13 (byte_count is an f32 there).
void format_file_size(char *buf, size_t cap, double byte_count);

}  // namespace et
