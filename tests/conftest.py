import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = ["test.txt", "nice.shakespeare.txt", "a_midsummer_nights_dream.txt"]  # test.zig:35-72


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


@pytest.fixture(scope="session")
def fixtures():
    return {n: open(os.path.join(GOLDEN, n), "rb").read() for n in FIXTURES}


@pytest.fixture(scope="session")
def golden_et():
    return {n: open(os.path.join(GOLDEN, n + ".et"), "rb").read() for n in FIXTURES}


@pytest.fixture(scope="session")
def codec():
    """One device context for the GPU tests; fails (not skips) when the CUDA path is unavailable."""
    import entreepy_b200 as et

    c = et.Codec(0)
    yield c
    c.close()


def make_cases(seed=7):
    """Small seeded inputs covering the shapes the codec must agree with the oracle on."""
    rng = np.random.default_rng(seed)
    cases = {}
    cases["one_byte"] = np.array([65], dtype=np.uint8)
    cases["single_symbol_run"] = np.full(1000, 7, dtype=np.uint8)  # root is a leaf: no entries, empty body
    cases["two_symbols"] = np.array([0, 1, 1, 0, 1], dtype=np.uint8)
    cases["nul_heavy"] = rng.choice(np.array([0, 0, 0, 1, 2], dtype=np.uint8), 5000)
    cases["all_256_once"] = np.arange(256, dtype=np.uint8)
    cases["all_256_uniform"] = rng.integers(0, 256, 70000, dtype=np.uint8)  # 256th symbol dropped (encode.zig:70)
    cases["uniform_255"] = rng.integers(1, 256, 70000, dtype=np.uint8)
    cases["dropped_symbol_dominates"] = np.concatenate([np.arange(256, dtype=np.uint8), np.full(20000, 255, np.uint8)])
    w = rng.random(256) ** 8
    cases["skewed_256"] = rng.choice(256, 100000, p=w / w.sum()).astype(np.uint8)
    fib = [1, 1]
    while len(fib) < 20:
        fib.append(fib[-1] + fib[-2])
    f = np.concatenate([np.full(c, s, np.uint8) for s, c in enumerate(fib)])
    rng.shuffle(f)
    cases["fibonacci_depth19"] = f
    for n in (15, 16, 17, 31, 4095, 4096, 4097, 8192 + 5, 65536 + 33):
        cases[f"text_len_{n}"] = rng.choice(np.frombuffer(b"etaoin shrdlu\n,.", dtype=np.uint8), n)
    return cases
