"""entreepy_b200 — B200-native Huffman compress/decompress path of typio/entreepy.

Host-side mirror of the reference's codec seam over the C ABI (include/entreepy_b200.h):

    encode(text, out_writer, flags)            <-> src/encode.zig:25
    decode(compressed_text, out_writer, flags) <-> src/decode.zig:13   (input is file[4..])
    EncodeFlags / DecodeFlags                  <-> encode.zig:9-14 / decode.zig:7-11

The compute runs in hand-written sm_100a CUDA kernels (entreepy_b200/csrc).  There is no CPU
implementation in this package; without the built library or without a B200 the calls raise.
"""
from .codec import (  # noqa: F401
    Codec,
    DecodeFlags,
    EncodeFlags,
    EntreepyError,
    build_codebook,
    decode,
    encode,
    parse_header,
    write_header,
)

__all__ = [
    "Codec", "EncodeFlags", "DecodeFlags", "EntreepyError", "encode", "decode",
    "build_codebook", "write_header", "parse_header",
]
