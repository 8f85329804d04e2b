"""One .et stream across the GPUs of a box: contiguous byte ranges, three tiny exchanges.

world == 1 goes straight to et_encode_dev / et_decode_dev.  world > 1 is built on the shard
entry points of the C ABI plus torch.distributed for the 2 KiB histogram all-reduce and the
bit-offset / symbol-count all-gathers; bulk data never leaves its GPU.
"""
from dataclasses import dataclass

import numpy as np

from . import _abi


class ShardPlan:
    """Rank r encodes input bytes [lo, hi): equal 16-byte-aligned slices, remainder to the last rank."""

    def __init__(self, n_total, world, rank):
        self.n_total, self.world, self.rank = int(n_total), int(world), int(rank)
        per = (self.n_total // self.world) & ~15
        self.bounds = [min(r * per, self.n_total) for r in range(self.world)] + [self.n_total]
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.hi - self.lo


@dataclass
class EncodeResult:
    total_bytes: int        # size of the whole .et file
    local_bytes: int        # bytes of it resident in this rank's output buffer
    header_bytes: int = 0
    body_bit_offset: int = 0  # global bit offset of this rank's first code inside the body
    body_bits: int = 0        # bits this rank produced


class ShardedCodec:
    def __init__(self, codec, plan, dist=None):
        self.codec, self.plan, self.dist = codec, plan, dist
        if plan.world > 1 and dist is None:
            raise ValueError("world > 1 needs torch.distributed")

    # ---- device-resident
    def encode(self, d_in, d_out, cap, flags, stream=None):
        if self.plan.world == 1:
            size = self.codec.encode_dev(d_in, self.plan.n_local, d_out, cap, flags, stream)
            return EncodeResult(total_bytes=size, local_bytes=size)
        raise NotImplementedError("sharded encode")

    def decode(self, res, d_et, d_out, cap, flags, stream=None):
        if self.plan.world == 1:
            return self.codec.decode_dev(d_et + 4, res.total_bytes - 4, d_out, cap, flags, stream)
        raise NotImplementedError("sharded decode")

    # ---- host buffers (the reference-facing calls)
    def encode_host(self, h_in, h_out, flags):
        if self.plan.world == 1:
            size = self.codec.encode_into(h_in, h_out, flags)
            return size, EncodeResult(total_bytes=size, local_bytes=size)
        raise NotImplementedError("sharded encode")

    def decode_host(self, res, h_et, h_out, flags):
        if self.plan.world == 1:
            return self.codec.decode_into(h_et[4 : res.total_bytes], h_out, flags)
        raise NotImplementedError("sharded decode")
