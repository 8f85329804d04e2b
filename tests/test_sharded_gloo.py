"""N>1 host logic of entreepy_b200.sharded over gloo on the CPU (world_size 2 and 3).

The three GPU calls (histogram, pack_shard, unpack_shard) are replaced by an oracle-based stand-in with the
same contracts, so what is tested here is everything around them: shard plan, histogram all-reduce, the
cross-rank scan of bit offsets, seam bytes, body redistribution, decode cuts with lead-in/look-ahead, the
entry/exit check and the repeat-from-the-true-boundary loop, output offsets and body_len clipping.
The GPU twin of this test (same protocol on real kernels) is in test_gpu_parity.py.
"""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class OracleBackend:
    """CPU stand-in for GpuBackend: numpy bit packing / bit-serial decoding with the same call contracts."""

    def __init__(self):
        from entreepy_b200 import _abi

        self.lib = _abi.load()

    def histogram(self, t_in, n):
        from oracle import oracle

        return oracle.histogram(t_in[:n].numpy())

    def shard_bits(self, counts, cb):
        import ctypes

        c = np.ascontiguousarray(counts, dtype=np.uint64)
        return int(self.lib.et_shard_bits(c.ctypes.data, ctypes.byref(cb)))

    def pack_shard(self, t_in, n, cb, phase, bits, t_out):
        lens = np.array([cb.code[s].length for s in range(256)], dtype=np.int64)
        codes = np.array([cb.code[s].data for s in range(256)], dtype=np.uint64)
        x = t_in[:n].numpy()
        l = lens[x]
        start = phase + np.concatenate([[0], np.cumsum(l)[:-1]]) if n else np.zeros(0, dtype=np.int64)
        assert int(l.sum()) == bits
        nbytes = (phase + bits + 7) // 8
        stream = np.zeros(nbytes * 8, dtype=np.uint8)
        c = codes[x]
        for k in range(int(lens.max()) if n else 0):
            m = l > k
            stream[start[m] + k] = ((c[m] >> (l[m] - 1 - k).astype(np.uint64)) & np.uint64(1)).astype(np.uint8)
        t_out[:nbytes] = torch.from_numpy(np.packbits(stream))
        return nbytes

    def unpack_shard(self, t_range, range_bytes, own_begin, own_end, dictionary, head_bit, t_out):
        table = {(int(dictionary.length[e]), int(dictionary.code[e])): int(dictionary.symbol[e])
                 for e in range(dictionary.n_entries)}
        bits = np.unpackbits(t_range[:range_bytes].numpy())
        end = range_bytes * 8

        def step(pos):
            v = 0
            for length in range(1, 33):
                if pos + length > end:
                    return None
                v = (v << 1) | int(bits[pos + length - 1])
                if (length, v) in table:
                    return length, table[(length, v)]
            raise AssertionError("not a code")

        if head_bit >= 0:
            pos = head_bit
        else:  # synchronise on the piece before the owned part, from a guess
            pos = own_begin * 8 - 128 if own_begin * 8 >= 128 else own_begin * 8
            while pos < own_begin * 8:
                pos += step(pos)[0]
        entry, n = pos, 0
        while pos < own_end * 8:
            got = step(pos)
            if got is None:
                break
            t_out[n] = got[1]
            n += 1
            pos += got[0]
        return n, entry, max(pos, own_end * 8) if pos >= own_end * 8 else own_end * 8

    def or_byte(self, t_buf, index, value):
        t_buf[index] |= value

    def first_byte(self, t_buf):
        return int(t_buf[0])


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from entreepy_b200 import sharded
        from oracle import oracle

        data = np.load(os.path.join(out_dir, f"{case}.npy"))
        plan = sharded.ShardPlan(data.size, world, rank)
        backend = HeadBackend() if case.endswith("_heads") else OracleBackend()
        coder = sharded.ShardedCodec(backend, plan, sharded.Comm(dist, torch.device("cpu")))
        t_in = torch.from_numpy(data[plan.lo:plan.hi].copy())
        t_body = torch.zeros(plan.n_local * 4 + 64, dtype=torch.uint8)
        res = coder.encode(t_in, t_body)

        # ---- the ranks' final bytes, concatenated, are the reference .et file
        want = oracle.encode(data, cap=9000 + 5 * data.size).tobytes()
        mine = t_body[res.own_lo - res.first_byte : res.own_hi - res.first_byte].numpy().tobytes()
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        assert res.header + b"".join(parts) == want, f"{case}: sharded .et differs from the oracle's"
        assert res.total_bytes == len(want)

        # ---- decode: redistribute the body, decode shards, place by offset
        t_range = coder.scatter_body(res, t_body)
        cuts, ranges = coder.decode_ranges(res.body_bytes)
        s, t = ranges[rank]
        assert t_range.numpy().tobytes() == want[len(res.header) + s : len(res.header) + t]
        t_out = torch.zeros(8 * max(t - s, 1) + 16, dtype=torch.uint8)
        dres = coder.decode(res.header[4:], res.body_bytes, t_range, t_out)
        texts = [None] * world
        dist.all_gather_object(texts, (dres.offset, t_out[: dres.n_local].numpy().tobytes(), dres.rounds))
        full = bytearray(data.size)
        for off, chunk, _ in texts:
            full[off : off + len(chunk)] = chunk
        lossless = np.unique(data).size < 256  # the reference drops a symbol when all 256 occur
        expect = data.tobytes() if lossless else oracle.decode(np.frombuffer(want, np.uint8)[4:], data.size).tobytes()
        assert sum(len(c) for _, c, _ in texts) == len(expect)
        assert bytes(full[: len(expect)]) == expect, f"{case}: sharded decode differs"
        if rank == 0:
            open(os.path.join(out_dir, f"{case}.rounds"), "w").write(str(max(r for _, _, r in texts)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cases():
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"etaoin shrdlu\n,.", dtype=np.uint8)
    return {
        "text_60k": rng.choice(alphabet, 60001),
        "text_tiny": rng.choice(alphabet, 700),
        "text_one_shard_worth": rng.choice(alphabet, 40),
        "uniform255_20k": rng.integers(1, 256, 20000, dtype=np.uint8),   # slow to synchronise: the repeat loop runs
        "all256_9k": rng.integers(0, 256, 9000, dtype=np.uint8),
        "two_symbols": rng.integers(0, 2, 5000, dtype=np.uint8),
        "text_30k_heads": rng.choice(alphabet, 30011),                     # seam bytes computed from gathered text heads
        "all256_9k_heads": rng.integers(0, 256, 9100, dtype=np.uint8),     # a dropped symbol may force the real exchange
        "text_tiny_heads": rng.choice(alphabet, 50),
    }


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_encode_decode_over_gloo(world):
    with tempfile.TemporaryDirectory() as d:
        cases = _cases()
        for name, data in cases.items():
            np.save(os.path.join(d, f"{name}.npy"), data)
        # world 3 repeats only the cases where a third rank changes the picture (keeps the CPU suite short)
        names = list(cases) if world == 2 else ["text_60k", "text_one_shard_worth", "uniform255_20k", "all256_9k_heads"]
        for name in names:
            mp.spawn(_worker, args=(world, _free_port(), name, d), nprocs=world, join=True)
        # text finds its boundaries at once; the 7/8-bit code of uniform bytes does not
        assert int(open(os.path.join(d, "text_60k.rounds")).read()) == 1
        assert int(open(os.path.join(d, "uniform255_20k.rounds")).read()) >= 2


def test_shard_plan_and_cuts():
    from entreepy_b200 import sharded

    for n, world in [(0, 2), (15, 2), (16, 3), (1000, 4), ((1 << 32) - 16, 8)]:
        plans = [sharded.ShardPlan(n, world, r) for r in range(world)]
        assert plans[0].lo == 0 and plans[-1].hi == n
        assert all(plans[r].hi == plans[r + 1].lo for r in range(world - 1))
        assert all(p.lo % 16 == 0 for p in plans)
    for body, world in [(0, 2), (31, 2), (32, 2), (100, 8), (4096, 3), (2516582400, 8)]:
        cuts = sharded.body_cuts(body, world)
        assert cuts[0] == 0 and cuts[-1] == body and cuts == sorted(cuts)
        assert all(c % 32 == 0 and (c + sharded.LOOK_AHEAD <= body or c == 0) for c in cuts[1:-1])


def test_first_output_byte_matches_a_packed_shard():
    """The seam byte a rank computes for its right neighbour equals what that neighbour's pack writes."""
    from entreepy_b200 import build_codebook, sharded
    from oracle import oracle

    rng = np.random.default_rng(9)
    be = OracleBackend()
    for trial in range(40):
        k = int(rng.integers(2, 200))
        data = rng.integers(0, k, 300, dtype=np.uint8)
        cb = build_codebook(oracle.histogram(data))
        shard = data[int(rng.integers(0, 200)):]
        phase = int(rng.integers(0, 8))
        bits = be.shard_bits(oracle.histogram(shard), cb)
        out = torch.zeros(shard.size * 4 + 8, dtype=torch.uint8)
        be.pack_shard(torch.from_numpy(shard.copy()), shard.size, cb, phase, bits, out)
        head = int.from_bytes(shard[:8].tobytes(), "little")
        got = sharded.first_output_byte(cb, head, shard.size, phase)
        if got is not None:
            assert got == int(out[0]), trial


class HeadBackend(OracleBackend):
    """Stand-in that also reports the shard's first text bytes, like GpuBackend (no third exchange)."""

    def head_symbols(self, t_in, n):
        k = min(n, 8)
        return int.from_bytes(bytes(t_in[:k].numpy()), "little") if k else 0
