"""One .et stream across the GPUs of a box: contiguous byte ranges, tiny exchanges.

Encode: rank r packs text bytes [lo_r, hi_r).  One exchange: an all-gather of the ranks' local histograms
(2 KiB each, with the first 8 text bytes of every shard).  Their sum is the global histogram — what
north_star's all-reduce delivers — and each of them, priced with the codebook, is that shard's bit count, so
the cross-GPU exclusive scan of bit offsets is computed on every rank without a second exchange.  Every rank
builds the same codebook (the host step is deterministic) and packs its slice directly at its final bit
position; bulk data never leaves its GPU.  (A third exchange of one byte per rank exists for shards too short
to fill their first byte.)

Decode: rank r decodes body bytes [B_r, B_{r+1}) (32-byte aligned cuts).  The stream has no index, so a
rank other than 0 does not know where its first codeword starts: it synchronises on the 64 bytes before
its range and reports the boundary it found; one all-gather of (symbols, entry, exit) per rank lets every
rank check entry_r == exit_{r-1}.  A rank whose guess was wrong (slowly synchronising codes) decodes again
from the true boundary; this repeats at most world-1 times.  Symbol counts give the output offsets.

The compute goes through a backend with three calls (histogram, pack_shard, unpack_shard): GpuBackend
binds them to the C ABI (et_histogram_dev, et_pack_shard_dev, et_unpack_shard_dev); the CPU tests plug in
an oracle-based stand-in to exercise this host logic over gloo.
"""
from dataclasses import dataclass, field

import ctypes
import numpy as np

from . import _abi
from .codec import build_codebook, parse_header, write_header

LEAD_IN = 64     # bytes before a decode shard used to find its first codeword boundary
LOOK_AHEAD = 32  # bytes after a decode shard (a codeword that begins inside it may end there)


class ShardPlan:
    """Rank r encodes input bytes [lo, hi): equal 16-byte-aligned slices, remainder to the last rank."""

    def __init__(self, n_total, world, rank):
        self.n_total, self.world, self.rank = int(n_total), int(world), int(rank)
        per = (self.n_total // self.world) & ~15
        self.bounds = [min(r * per, self.n_total) for r in range(self.world)] + [self.n_total]
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]
        self.n_local = self.hi - self.lo


def body_cuts(body_bytes, world):
    """Decode shard boundaries B_0..B_world: equal parts of the body, cut on 32-byte sectors.  Every cut but
    the last leaves at least LOOK_AHEAD bytes after it (short bodies give the later ranks nothing)."""
    last_cut = ((body_bytes - LOOK_AHEAD) & ~31) if body_bytes >= LOOK_AHEAD else 0
    cuts = [0]
    for r in range(1, world):
        cuts.append(max(cuts[-1], min(last_cut, (r * body_bytes // world) & ~31)))
    cuts.append(body_bytes)
    return cuts


@dataclass
class EncodeResult:
    total_bytes: int                 # size of the whole .et file
    header: bytes = b""              # the same on every rank
    bit_offsets: list = field(default_factory=list)  # [world + 1] global bit offset of each rank's first code in the body
    first_byte: int = 0              # body byte index of this rank's out[0]
    local_bytes: int = 0             # bytes this rank wrote (out[0 : local_bytes])
    own_lo: int = 0                  # body bytes [own_lo, own_hi) are final in this rank's buffer
    own_hi: int = 0
    codebook: object = None

    @property
    def body_bytes(self):
        return (self.bit_offsets[-1] + 7) // 8


@dataclass
class DecodeResult:
    n_local: int      # valid symbols in this rank's output buffer
    offset: int       # position of the first of them in the text
    rounds: int = 1   # 1 = every rank's guessed entry was right


class Comm:
    """The three exchanges, over torch.distributed (nccl: tensors on the GPU, gloo: on the CPU)."""

    def __init__(self, dist, device):
        self.dist, self.device = dist, device
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0

    def allreduce_counts(self, counts):
        if self.world == 1:
            return counts
        import torch

        t = torch.from_numpy(counts.astype(np.int64)).to(self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy().astype(np.uint64)

    def allgather_ints(self, values):
        """values: list of python ints (non-negative, < 2^63) -> [world][len(values)]."""
        if self.world == 1:
            return [list(values)]
        return self.allgather_array(np.asarray(values, dtype=np.int64)).tolist()

    def allgather_array(self, values):
        """values: int64[k] (non-negative) -> int64[world, k]."""
        import torch

        k = int(values.size)
        if self.world == 1:
            return values.reshape(1, k).copy()
        if self.device.type != "cuda":
            t = torch.from_numpy(np.ascontiguousarray(values, dtype=np.int64)).to(self.device)
            out = torch.empty(self.world * k, dtype=torch.int64, device=self.device)
            self.dist.all_gather_into_tensor(out, t)
            return out.cpu().view(self.world, k).numpy().copy()
        # GPU: page-locked staging and device buffers are allocated once; the exchanges are a few hundred bytes
        # and their cost is launch and copy latency, not bandwidth
        if getattr(self, "_cap", 0) < k:
            self._cap = max(512, k)
            self._h_send = torch.empty(self._cap, dtype=torch.int64).pin_memory()
            self._h_recv = torch.empty(self._cap * self.world, dtype=torch.int64).pin_memory()
            self._d_send = torch.empty(self._cap, dtype=torch.int64, device=self.device)
            self._d_recv = torch.empty(self._cap * self.world, dtype=torch.int64, device=self.device)
        self._h_send.numpy()[:k] = values
        self._d_send[:k].copy_(self._h_send[:k], non_blocking=True)
        self.dist.all_gather_into_tensor(self._d_recv[: self.world * k], self._d_send[:k])
        self._h_recv[: self.world * k].copy_(self._d_recv[: self.world * k], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._h_recv.numpy()[: self.world * k].reshape(self.world, k).copy()

    def all_to_all_bytes(self, send, send_splits, recv_splits):
        """Uneven all-to-all of byte ranges (point-to-point pairs: works on nccl and gloo alike)."""
        import torch

        recv = torch.empty(int(sum(recv_splits)), dtype=torch.uint8, device=send.device)
        ops, so, ro = [], 0, 0
        for q in range(self.world):
            ns, nr = int(send_splits[q]), int(recv_splits[q])
            if q == self.rank:
                if ns:
                    recv[ro : ro + nr].copy_(send[so : so + ns])
            else:
                if ns:
                    ops.append(self.dist.P2POp(self.dist.isend, send[so : so + ns], q))
                if nr:
                    ops.append(self.dist.P2POp(self.dist.irecv, recv[ro : ro + nr], q))
            so, ro = so + ns, ro + nr
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        return recv


class GpuBackend:
    """histogram / pack_shard / unpack_shard on CUDA tensors through the C ABI."""

    def __init__(self, codec, stream=None):
        self.codec, self.stream = codec, stream

    def histogram(self, t_in, n):
        counts = self.codec.histogram_dev(t_in.data_ptr(), n, self.stream)
        self.hist_ms = self.codec.last_stage_ms()[0]  # et_histogram_dev times its own kernel + read-back
        return counts

    def shard_bits(self, counts, cb):
        return self.codec.shard_bits(counts, cb)

    def pack_shard(self, t_in, n, cb, phase, bits, t_out):
        return self.codec.pack_shard_dev(t_in.data_ptr(), n, cb, phase, bits, t_out.data_ptr(), t_out.numel(), self.stream)

    def unpack_shard(self, t_range, range_bytes, own_begin, own_end, dictionary, head_bit, t_out):
        return self.codec.unpack_shard_dev(t_range.data_ptr(), range_bytes, own_begin, own_end, dictionary, head_bit,
                                           t_out.data_ptr(), t_out.numel(), self.stream)

    def or_byte(self, t_buf, index, value):
        if value:
            t_buf[index] |= value

    def first_byte(self, t_buf):
        return int(t_buf[0].item())

    def head_symbols(self, t_in, n):
        """The first (up to 8) text bytes of the shard, as an int (little endian)."""
        k = min(n, 8)
        return int.from_bytes(bytes(t_in[:k].cpu().numpy()), "little") if k else 0

    def head_symbols_begin(self, t_in, n):
        """Starts the copy of the first text bytes into page-locked memory; head_symbols_end() reads them after
        the next synchronisation of the stream (the histogram's), so the copy costs no round trip of its own."""
        import torch

        k = min(n, 8)
        if not hasattr(self, "_head"):
            self._head = torch.zeros(8, dtype=torch.uint8).pin_memory()
            self._head_done = torch.cuda.Event()
        if k:
            self._head[:k].copy_(t_in[:k], non_blocking=True)
            self._head_done.record()
        return k

    def head_symbols_end(self, k):
        if k:
            self._head_done.synchronize()  # already signalled when the histogram ran on the same stream
        return int.from_bytes(bytes(self._head[:k].numpy()), "little") if k else 0


def first_output_byte(cb, head, n, phase):
    """First byte a shard writes when packed at bit `phase`: its leading codes behind `phase` zero bits.
    None when the shard's first 8 symbols do not fill the byte (then the ranks exchange the real byte)."""
    acc, nbits = 0, 0
    for i in range(min(n, 8)):
        c = cb.code[(head >> (8 * i)) & 0xFF]
        length = int(c.length)
        if length > 32:
            return None
        acc = (acc << length) | (int(c.data) & ((1 << length) - 1))
        nbits += length
        if nbits >= 8 - phase:
            return (acc >> (nbits - (8 - phase))) & 0xFF
    return None


class ShardedCodec:
    def __init__(self, backend, plan, comm):
        self.backend, self.plan, self.comm = backend, plan, comm

    # ------------------------------------------------------------------ encode
    def encode(self, t_in, t_out):
        """t_in: this rank's slice of the text; t_out: room for its part of the body (n_local + 16 bytes is
        what the reference's 7200+n scratch bound guarantees).  Returns EncodeResult; the .et file is
        header + the ranks' bodies[own_lo:own_hi] in rank order."""
        p, be = self.plan, self.backend
        pending = be.head_symbols_begin(t_in, p.n_local) if hasattr(be, "head_symbols_begin") else None
        local = be.histogram(t_in, p.n_local)                       # K1 (synchronises the stream)
        # ONE exchange: every rank's local histogram (2 KiB each).  The sum is the global histogram (what an
        # all-reduce would give); the local ones, priced with the codebook, are every shard's bit count — the
        # cross-GPU scan of bit offsets needs no second exchange.  The first text bytes of every shard ride
        # along so that each rank can work out the seam byte of its right neighbours.
        if pending is not None:
            head = be.head_symbols_end(pending)
        else:
            head = be.head_symbols(t_in, p.n_local) if hasattr(be, "head_symbols") else 0
        send = np.empty(259, dtype=np.int64)
        send[:256] = local
        send[256:] = (head & 0x7FFFFFFFFFFFFFFF, head >> 63, p.n_local)
        packed = self.comm.allgather_array(send)                    # [world, 259]
        locals_ = packed[:, :256].astype(np.uint64)
        counts = locals_.sum(axis=0, dtype=np.uint64)
        cb = build_codebook(counts)                                 # same tables on every rank
        header = write_header(cb, p.n_total)
        # bits of every shard = its local counts priced with the code lengths (what et_shard_bits computes)
        lengths = np.frombuffer(cb, dtype=np.dtype([("data", "<u4"), ("length", "u1"), ("pad", "V3")]), count=256)["length"]
        shard_bits = [int(x) for x in (locals_ * lengths.astype(np.uint64)).sum(axis=1, dtype=np.uint64)]
        bits = shard_bits[p.rank]
        tails = packed[:, 256:].tolist()
        gathered = [[shard_bits[r]] + tails[r] for r in range(p.world)]
        offs = [0]
        for r in range(p.world):
            offs.append(offs[-1] + gathered[r][0])
        my_off = offs[p.rank]
        nbytes = be.pack_shard(t_in, p.n_local, cb, my_off & 7, bits, t_out)   # K2 at the final bit position
        res = EncodeResult(total_bytes=len(header) + (offs[-1] + 7) // 8, header=header, bit_offsets=offs,
                           first_byte=my_off >> 3, local_bytes=nbytes, codebook=cb)
        # the byte in which rank r ends may also hold the first bits of the ranks after it
        if p.world > 1:
            last = (offs[p.rank + 1] - 1) >> 3 if bits else -1
            sharers = [q for q in range(p.rank + 1, p.world)
                       if offs[q + 1] > offs[q] and (offs[q] & 7) and (offs[q] >> 3) == last]
            firsts = {}
            for q in sharers:
                hq = gathered[q][1] | (gathered[q][2] << 63)
                firsts[q] = first_output_byte(cb, hq, gathered[q][3], offs[q] & 7) if hasattr(be, "head_symbols") else None
            # every rank takes the same decision: the inputs are the gathered values
            need_exchange = False
            for q in range(1, p.world):
                if offs[q + 1] > offs[q] and (offs[q] & 7):
                    hq = gathered[q][1] | (gathered[q][2] << 63)
                    if not hasattr(be, "head_symbols") or first_output_byte(cb, hq, gathered[q][3], offs[q] & 7) is None:
                        need_exchange = True
            if need_exchange:  # exchange 3 (tiny shards, dropped symbols): the real first bytes
                got = self.comm.allgather_ints([be.first_byte(t_out) if nbytes else 0])
                firsts = {q: got[q][0] for q in sharers}
            merged = 0
            for q in sharers:
                merged |= firsts[q]
            if merged:
                be.or_byte(t_out, last - res.first_byte, merged)
        # bytes of the body that are final in this rank's buffer: a shard that starts inside a byte leaves
        # that byte to the rank that started it
        lo = (my_off + 7) >> 3 if p.rank > 0 else 0
        hi = (offs[p.rank + 1] + 7) >> 3
        res.own_lo, res.own_hi = min(lo, hi), hi
        return res

    # ------------------------------------------------------------------ body redistribution (setup for decode)
    def decode_ranges(self, body_bytes):
        cached = getattr(self, "_ranges_of", None)
        if cached is not None and cached[0] == body_bytes:
            return cached[1], cached[2]
        cuts = body_cuts(body_bytes, self.plan.world)
        ranges = []
        for r in range(self.plan.world):
            s = max(cuts[r] - LEAD_IN, 0) if r > 0 else 0
            t = min(cuts[r + 1] + LOOK_AHEAD, body_bytes) if r + 1 < self.plan.world else body_bytes
            ranges.append((s, t))
        self._ranges_of = (body_bytes, cuts, ranges)
        return cuts, ranges

    def scatter_body(self, res, t_body):
        """From the layout encode() leaves (rank q holds body bytes [own_lo, own_hi)) to the layout decode()
        wants (rank r holds [S_r, T_r): its equal share plus lead-in and look-ahead).  One all-to-all; this is
        the job a file reader does when the stream comes from disk, so it is not part of the decode."""
        import torch

        p = self.plan
        cuts, ranges = self.decode_ranges(res.body_bytes)
        if p.world == 1:
            return t_body[: res.body_bytes]
        owns = self.comm.allgather_ints([res.own_lo, res.own_hi])
        send_parts, send_splits, recv_splits = [], [], []
        for r in range(p.world):
            lo, hi = max(res.own_lo, ranges[r][0]), min(res.own_hi, ranges[r][1])
            n = max(hi - lo, 0)
            send_splits.append(n)
            if n:
                send_parts.append(t_body[lo - res.first_byte : hi - res.first_byte])
            s, t = ranges[p.rank]
            recv_splits.append(max(min(owns[r][1], t) - max(owns[r][0], s), 0))
        send = torch.cat(send_parts) if send_parts else torch.empty(0, dtype=torch.uint8, device=t_body.device)
        return self.comm.all_to_all_bytes(send, send_splits, recv_splits)

    # ------------------------------------------------------------------ decode
    def decode(self, header, body_bytes, t_range, t_out):
        """header: the .et bytes after the magic up to the body (every rank reads them); t_range: this rank's
        range of the body as decode_ranges() lays it out; t_out: room for its text (any rank may get up to
        8 symbols per byte in theory; the bench sizes it from the plan)."""
        p, be = self.plan, self.backend
        dictionary = parse_header(header)
        cuts, ranges = self.decode_ranges(body_bytes)
        s, t = ranges[p.rank]
        own_begin, own_end = cuts[p.rank] - s, cuts[p.rank + 1] - s
        head_bit = 0 if cuts[p.rank] == 0 else -1  # a share that begins with the body begins on a codeword
        rounds, redo = 0, True
        n = entry = exit_ = 0
        while True:
            if redo and own_begin == own_end:  # an empty share (short bodies give the later cuts nothing): no kernel to run
                n, entry, exit_ = 0, own_begin * 8, own_end * 8
            elif redo:
                n, entry, exit_ = be.unpack_shard(t_range, t - s, own_begin, own_end, dictionary, head_bit, t_out)   # K3-K5
            rounds += 1
            if p.world == 1:
                return DecodeResult(n_local=min(n, int(dictionary.body_len)), offset=0, rounds=rounds)
            # exchange: symbols, where my first codeword began, where my last one ended (bits past the cut)
            info = self.comm.allgather_ints([n, entry - own_begin * 8, exit_ - own_end * 8])
            wrong, prev_exit = [], 0  # rank 0 starts on bit 0 of the body
            for r in range(p.world):
                if cuts[r] == cuts[r + 1]:
                    continue  # an empty share passes its neighbour's end on
                if r > 0 and info[r][1] != prev_exit:
                    wrong.append((r, prev_exit))
                prev_exit = info[r][2]
            if not wrong:
                break
            mine = [e for r, e in wrong if r == p.rank]
            redo = bool(mine)
            if redo:
                head_bit = own_begin * 8 + mine[0]
            if rounds > p.world + 1:
                raise RuntimeError("sharded decode did not settle")  # cannot happen: rank r is right after r rounds
        offset = sum(info[r][0] for r in range(p.rank))
        valid = max(min(info[p.rank][0], int(dictionary.body_len) - offset), 0)
        return DecodeResult(n_local=valid, offset=offset, rounds=rounds)


class NativeShardedCodec(ShardedCodec):
    """The same protocol with the exchanges INSIDE the library: et_encode_sharded_dev / et_decode_sharded_dev run the
    kernels, the all-gathers (NCCL on the context's stream, or a host callback) and the host steps between them, so a
    step costs one call per direction and no Python between kernel and collective.  `comm` is an et_comm made by
    Codec.comm_nccl / Codec.comm_callback; `py_comm` (a Comm) is only used by scatter_body, the file reader's job."""

    def __init__(self, codec, plan, comm, py_comm, stream=None):
        super().__init__(GpuBackend(codec, stream), plan, py_comm)
        self.codec, self.ncomm, self.stream = codec, comm, stream

    def encode(self, t_in, t_out):
        p = self.plan
        r = self.codec.encode_sharded_dev(self.ncomm, t_in.data_ptr(), p.n_local, t_out.data_ptr(), t_out.numel(), 0, self.stream)
        self.backend.hist_ms = self.codec.last_stage_ms()[0]
        # (string_at: slicing the ctypes array would build a Python list of a thousand ints on every call)
        res = EncodeResult(total_bytes=int(r.total_bytes), header=ctypes.string_at(ctypes.addressof(r.header), int(r.header_len)), bit_offsets=[0, int(r.body_bytes) * 8],
                           first_byte=int(r.first_byte), local_bytes=int(r.local_bytes), own_lo=int(r.own_lo), own_hi=int(r.own_hi))
        res.bit_offset = int(r.bit_offset)
        res.n_total = int(r.n_total)
        return res

    def decode(self, header, body_bytes, t_range, t_out):
        p = self.plan
        cuts, ranges = self.decode_ranges(body_bytes)
        s, t = ranges[p.rank]
        own_begin, own_end = cuts[p.rank] - s, cuts[p.rank + 1] - s
        r = self.codec.decode_sharded_dev(self.ncomm, header, t_range.data_ptr(), t - s, own_begin, own_end, cuts[p.rank] == 0,
                                          t_out.data_ptr(), t_out.numel(), 0, self.stream)
        return DecodeResult(n_local=int(r.n_local), offset=int(r.offset), rounds=int(r.rounds))
