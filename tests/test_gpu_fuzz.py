"""Seeded random inputs through the whole CUDA path (lane-run pack, lane-interleaved decoder forced on by
et_ctx_set_tuning(ET_TUNE_LANE_MIN_BYTES, 0)) against the oracle: random alphabets and skews, sizes around the region sizes (2048 symbols
on encode, 4224 stream bytes on decode), every alignment of input, stream and output.  Bar: bit-exact."""
import os

import numpy as np
import pytest

import entreepy_b200 as et
from entreepy_b200 import _abi
from oracle import oracle

pytestmark = pytest.mark.gpu


def _case(rng):
    m = int(rng.integers(2, 257))                       # alphabet size
    skew = float(rng.choice([0.0, 0.5, 1.0, 2.0, 4.0, 8.0]))
    w = rng.random(m) ** skew if skew else np.ones(m)
    syms = rng.permutation(256)[:m].astype(np.uint8)
    n = int(rng.choice([rng.integers(1, 5000), rng.integers(5000, 70000), rng.integers(70000, 600000),
                        rng.integers(600000, 3000000)]))
    data = syms[rng.choice(m, n, p=w / w.sum())]
    return data


@pytest.fixture()
def lanes(codec):
    codec.set_tuning(et._abi.TUNE_LANE_MIN_BYTES, 0)
    yield
    codec.set_tuning(et._abi.TUNE_LANE_MIN_BYTES, -1)


def test_fuzz_encode_decode_against_oracle(codec, lanes):
    import torch

    rng = np.random.default_rng(20261018)
    done = 0
    cases = int(os.environ.get("ET_FUZZ_CASES", "60"))
    for it in range(cases):
        data = _case(rng)
        if np.unique(data).size < 2:
            continue
        want = oracle.encode(data, cap=9000 + 5 * data.size).tobytes()
        in_off, out_off, dec_off = (int(x) for x in rng.integers(0, 16, 3))
        d_in = torch.zeros(data.size + 32, dtype=torch.uint8, device="cuda")
        d_in[in_off : in_off + data.size] = torch.from_numpy(data).cuda()
        d_et = torch.zeros(len(want) + 64, dtype=torch.uint8, device="cuda")
        size = codec.encode_dev(d_in.data_ptr() + in_off, data.size, d_et.data_ptr() + out_off, len(want) + 16,
                                _abi.FLAG_WRITE_OUTPUT | _abi.FLAG_NO_SCRATCH_LIMIT)
        assert size == len(want), (it, data.size)
        got = d_et[out_off : out_off + size].cpu().numpy().tobytes()
        assert got == want, (it, data.size, in_off, out_off)
        assert int(d_et[:out_off].sum()) == 0 and int(d_et[out_off + size :].sum()) == 0  # nothing outside the file
        text = oracle.decode(want[4:], data.size).tobytes()  # what the stream holds (a dropped 256th symbol is gone)
        d_out = torch.zeros(data.size + 32, dtype=torch.uint8, device="cuda")
        m = codec.decode_dev(d_et.data_ptr() + out_off + 4, size - 4, d_out.data_ptr() + dec_off, data.size)
        assert m == len(text), (it, data.size)
        assert d_out[dec_off : dec_off + m].cpu().numpy().tobytes() == text, (it, data.size, out_off, dec_off)
        assert int(d_out[:dec_off].sum()) == 0 and int(d_out[dec_off + m :].sum()) == 0
        done += 1
    assert done >= cases * 3 // 4
