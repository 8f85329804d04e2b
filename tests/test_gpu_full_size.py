"""BASELINE.json's configurations at their full sizes, through properties that do not need the oracle to
produce gigabytes: the .et size is header + ceil(sum(count x length) / 8) for the codebook the ORACLE builds
from the same histogram, the header bytes are the oracle's, decode(encode(x)) == x; every configuration up to
1 GiB is in addition compared byte for byte with the oracle's .et (seconds of host time each; text-4G, which the
oracle would take a minute on, is covered by the size/header/round-trip properties in bench.py's configs block).
"""
import hashlib
import json
import os

import numpy as np
import pytest

import entreepy_b200 as et
from conftest import GOLDEN
from entreepy_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.gpu

MAN = json.load(open(os.path.join(GOLDEN, "manifest.json")))
CONFIGS = {  # name -> (bytes, weights, compare the whole .et with the oracle)
    "text-5M": (5452595, synth.text_weights(MAN["midsummer_histogram"]), True),   # BASELINE config 2
    "text-1G": (1 << 30, synth.text_weights(MAN["midsummer_histogram"]), True),   # ~12 s of oracle for the byte compare
    "uniform255-256M": (1 << 28, synth.uniform_weights(1), True),
    "uniform256-256M": (1 << 28, synth.uniform_weights(0), True),
    "fib32-256M": (1 << 28, synth.fibonacci_weights(32), True),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_size_config(codec, name):
    import torch

    n, weights, whole = CONFIGS[name]
    if name.startswith("fib32"):  # exact counts: i.i.d. sampling does not give a chain of depth 32
        text = synth.shuffled_dev(synth.fibonacci_counts(n, 32))
        counts = codec.histogram_dev(text.data_ptr(), n)
        assert [int(x) for x in counts] == synth.fibonacci_counts(n, 32)
    else:
        thr = synth.thresholds_from_weights(weights)
        text = torch.empty(n, dtype=torch.uint8, device="cuda")
        codec.synth_dev(text.data_ptr(), n, synth.SEED, 0, thr)
        counts = codec.histogram_dev(text.data_ptr(), n)
    assert int(counts.sum()) == n
    # the codebook: ours and the oracle's from the same counts
    cb = et.build_codebook(counts)
    o_code, o_len = oracle.build_dictionary(counts)
    assert [cb.code[s].length for s in range(256)] == [int(x) for x in o_len]
    assert [cb.code[s].data & 0xFFFFFFFF for s in range(256)] == [int(x) for x in o_code]
    header = et.write_header(cb, n)
    bits = int((counts.astype(object) * o_len.astype(object)).sum())
    if name.startswith("fib32"):
        assert int(o_len.max()) == 32  # a chain: the deepest codes the reference represents faithfully
    enc = torch.empty(n + n // 4 + 65536, dtype=torch.uint8, device="cuda")
    size = codec.encode_dev(text.data_ptr(), n, enc.data_ptr(), enc.numel(), et._abi.FLAG_WRITE_OUTPUT | et._abi.FLAG_NO_SCRATCH_LIMIT)
    assert size == len(header) + (bits + 7) // 8
    assert enc[: len(header)].cpu().numpy().tobytes() == header
    if whole:
        want = oracle.encode(text.cpu().numpy(), cap=n + n // 4 + 65536)
        assert size == want.size
        assert hashlib.sha256(enc[:size].cpu().numpy().tobytes()).hexdigest() == hashlib.sha256(want.tobytes()).hexdigest()
    dec = torch.zeros(n, dtype=torch.uint8, device="cuda")
    got = codec.decode_dev(enc.data_ptr() + 4, size - 4, dec.data_ptr(), n)
    dropped = [s for s in range(256) if counts[s] and cb.code[s].length == 0]
    if dropped:  # all 256 byte values occur: the reference gives the last symbol in sort order no code (encode.zig:70)
        assert name.startswith("uniform256") and len(dropped) == 1
        keep = text[text != dropped[0]]
        assert got == keep.numel() and torch.equal(dec[:got], keep)
    else:
        assert got == n and torch.equal(dec, text)
