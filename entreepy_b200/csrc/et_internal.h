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
    uint32_t narrow[256];  // (emitted << 6) | length, valid when max_length <= kNarrowMaxLen
    uint64_t wide_code[256];
    uint8_t wide_len[256];
    uint32_t max_length;
    bool narrow_ok;
};
constexpr uint32_t kNarrowMaxLen = 26;

int make_pack_tables(const et_codebook &cb, PackTables *t);

// ---------------------------------------------------------------- decoder tables (host -> device)
// First level: kLutBits-bit window -> packed entry.
//   [ 7: 0] sym0      first symbol in the window
//   [15: 8] sym1      second symbol (valid when len01 != 0)
//   [19:16] len0      length of the first code, 1..12; 0 = code longer than the window
//                     (or no code at all): then [15:0] is the trie node reached, 0xFFFF = invalid
//   [23:20] len01     bits consumed by the first two codes, 0 = second does not fit
//   [27:24] bits_all  bits consumed by every whole code that fits in the window
//   [31:28] cnt_all   how many codes that is (1..12)
// Second level: binary trie, node = (child1 << 16) | child0; child < 0x8000 = node index,
// 0x8000|sym = leaf, 0xFFFF = no such code.
constexpr int kLutBits = 12;
constexpr int kLutSize = 1 << kLutBits;
constexpr uint32_t kMaxTrieNodes = 8192;
constexpr uint32_t kChildLeaf = 0x8000u;
constexpr uint32_t kChildNone = 0xFFFFu;

struct UnpackTables {
    uint32_t lut[kLutSize];
    uint32_t nodes[kMaxTrieNodes];
    uint32_t n_nodes;
    uint32_t max_length;
    uint32_t min_length;
    bool complete;  // every window decodes (Kraft sum == 1)
};

int make_unpack_tables(const et_dictionary &dict, UnpackTables *t);

// "{d} B" / "{d:.2} KB" ... of utils.zig:3-13 (byte_count is an f32 there).
void format_file_size(char *buf, size_t cap, double byte_count);

}  // namespace et
