"""The lane-interleaved decoder (regions of 32 chunks per warp) on inputs of every shape.

Long streams take it by default; et_ctx_set_tuning(ET_TUNE_LANE_MIN_BYTES, 0) sends short ones through it as well, so the
oracle can check it at sizes it finishes in seconds.  Bar: bit-exact (the original text).
"""
import numpy as np
import pytest

import entreepy_b200 as et
from conftest import make_cases
from entreepy_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


def _oracle_et(data):
    return oracle.encode(data, cap=9000 + 5 * int(np.asarray(data).size)).tobytes()


@pytest.fixture()
def lanes(codec):
    codec.set_tuning(et._abi.TUNE_LANE_MIN_BYTES, 0)
    yield
    codec.set_tuning(et._abi.TUNE_LANE_MIN_BYTES, -1)


def test_lane_decoder_on_the_oracle_cases(codec, lanes):
    for name, data in make_cases().items():
        if name in ("one_byte", "single_symbol_run"):
            continue
        stream = _oracle_et(data)[4:]
        want = oracle.decode(stream, data.size).tobytes()
        n, out = codec.decode(stream)
        assert out.tobytes() == want, name


def test_lane_decoder_text_sizes(codec, lanes, manifest):
    thr = synth.thresholds_from_weights(synth.text_weights(manifest["midsummer_histogram"]))
    host = synth.generate((3 << 20) + 321, thr)
    # region = 4224 body bytes ~ 7200 symbols: streams of 2..3 regions, ragged ends, a few hundred regions
    for n in (14000, 14500, 15000, 21599, 21600, 21700, 30011, 100003, 1 << 20, host.size):
        data = host[:n]
        m, out = codec.decode(_oracle_et(data)[4:])
        assert m == n and out.tobytes() == data.tobytes(), n
        assert codec.last_decode_rounds == 2


def test_lane_decoder_alignments_and_capacity(codec, lanes, manifest):
    import torch

    thr = synth.thresholds_from_weights(synth.text_weights(manifest["midsummer_histogram"]))
    data = synth.generate(300007, thr, seed=11)
    et_file = _oracle_et(data)
    src = torch.from_numpy(np.frombuffer(et_file, dtype=np.uint8).copy()).cuda()
    dev = torch.zeros(len(et_file) + 64, dtype=torch.uint8, device="cuda")
    out = torch.zeros(data.size + 64, dtype=torch.uint8, device="cuda")
    for phase in range(16):
        dev[phase : phase + len(et_file)] = src
        o = (phase * 7) % 16
        out.zero_()
        got = codec.decode_dev(dev.data_ptr() + phase + 4, len(et_file) - 4, out.data_ptr() + o, data.size)
        assert got == data.size and out[o : o + got].cpu().numpy().tobytes() == data.tobytes(), phase
        assert int(out[o + got :].sum()) == 0 and int(out[:o].sum()) == 0  # nothing outside the text
    # output clipped by the caller's capacity: exactly cap bytes are written, NoSpaceLeft reported
    for cap in (1, 4223, 7200, 100000, data.size - 1):
        out.zero_()
        with pytest.raises(et.EntreepyError) as e:
            codec.decode_dev(dev.data_ptr() + 15 + 4, len(et_file) - 4, out.data_ptr(), cap)
        assert e.value.name == "NoSpaceLeft"
        assert out[:cap].cpu().numpy().tobytes() == data[:cap].tobytes() and int(out[cap:].sum()) == 0, cap


def test_lane_decoder_long_codes_and_skew(codec, lanes):
    # Fibonacci weights: one-bit codes next to codes of 27 and 32 bits (the trie path inside the flat loop);
    # the text of a region is up to 8x its stream bytes, so the stage of a warp is sized per stream
    for depth in (19, 27, 32):
        w = synth.fibonacci_weights(depth)
        data = synth.generate(1 << 21, synth.thresholds_from_weights(w), seed=depth)
        data[: depth + 1] = np.arange(depth + 1, dtype=np.uint8)
        m, out = codec.decode(_oracle_et(data)[4:])
        assert m == data.size and out.tobytes() == data.tobytes(), depth
    rng = np.random.default_rng(17)
    w = rng.random(256) ** 8
    data = rng.choice(256, 1 << 21, p=w / w.sum()).astype(np.uint8)
    data[:256] = np.arange(256, dtype=np.uint8)
    stream = _oracle_et(data)[4:]
    m, out = codec.decode(stream)
    assert out.tobytes() == oracle.decode(stream, data.size).tobytes()
    # two symbols, one bit each: eight symbols per stream byte
    data = rng.integers(0, 2, 1 << 20, dtype=np.uint8)
    m, out = codec.decode(_oracle_et(data)[4:])
    assert m == data.size and out.tobytes() == data.tobytes()


def test_lane_decoder_repairs_wrong_guesses(codec, lanes):
    # 40 of 48 symbols equiprobable: most codes have 6 or 7 bits and a wrong parse survives for a long time, so
    # the run-up guesses are often wrong and the repair rounds have to settle the entries
    rng = np.random.default_rng(5)
    w = np.ones(48)
    w[:8] = 9.0  # lengths 3..7: a spread wide enough for the lane path, still slow to synchronise
    data = rng.choice(48, 1 << 20, p=w / w.sum()).astype(np.uint8)
    lengths = oracle.build_dictionary(oracle.histogram(data))[1]
    assert int(lengths.max()) - int(lengths[lengths > 0].min()) > 2
    m, out = codec.decode(_oracle_et(data)[4:])
    assert m == data.size and out.tobytes() == data.tobytes()
