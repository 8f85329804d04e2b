"""Step-by-step GPU bring-up check with first-difference diagnostics (developer tool, run under gpurun)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import entreepy_b200 as et  # noqa: E402
from entreepy_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402


def first_diff(a, b):
    a = np.frombuffer(a, dtype=np.uint8)
    b = np.frombuffer(b, dtype=np.uint8)
    n = min(a.size, b.size)
    d = np.nonzero(a[:n] != b[:n])[0]
    if d.size == 0:
        return None if a.size == b.size else n
    return int(d[0])


def check(name, got, want):
    fd = first_diff(got, want)
    if fd is None:
        print(f"  ok   {name} ({len(want)} B)")
        return True
    nd = int((np.frombuffer(got, np.uint8)[: min(len(got), len(want))] != np.frombuffer(want, np.uint8)[: min(len(got), len(want))]).sum())
    print(f"  FAIL {name}: len got={len(got)} want={len(want)} first diff @ {fd} ndiff={nd}")
    print("       got ", bytes(got[max(0, fd - 4) : fd + 12]).hex(" "))
    print("       want", bytes(want[max(0, fd - 4) : fd + 12]).hex(" "))
    return False


def main():
    import json

    man = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
    thr = synth.thresholds_from_weights(synth.text_weights(man["midsummer_histogram"]))
    codec = et.Codec(0)
    ok = True
    cases = {n: open(os.path.join(ROOT, "tests/golden", n), "rb").read() for n in ("test.txt", "nice.shakespeare.txt", "a_midsummer_nights_dream.txt")}
    for n in (4096, 4097, 100000, 1 << 20, (1 << 22) + 5):
        cases[f"text_{n}"] = synth.generate(n, thr).tobytes()
    rng = np.random.default_rng(1)
    cases["uniform255_64k"] = rng.integers(1, 256, 65536, dtype=np.uint8).tobytes()
    cases["uniform256_64k"] = rng.integers(0, 256, 65536, dtype=np.uint8).tobytes()
    for name, data in cases.items():
        print(name, len(data))
        h = codec.histogram(data)
        if not np.array_equal(h, oracle.histogram(data)):
            print("  FAIL histogram", int((h != oracle.histogram(data)).sum()), "bins differ; sum", int(h.sum()))
            ok = False
        want = oracle.encode(data, cap=9000 + 5 * len(data)).tobytes()
        try:
            n, enc = codec.encode(data, et.EncodeFlags(write_output=True, no_scratch_limit=True))
            ok &= check("encode", enc.tobytes(), want)
        except et.EntreepyError as e:
            print("  FAIL encode raised", e)
            ok = False
        try:
            t0 = time.time()
            m, dec = codec.decode(want[4:])
            wd = oracle.decode(want[4:], len(data)).tobytes()
            ok &= check("decode", dec.tobytes(), wd)
        except et.EntreepyError as e:
            print("  FAIL decode raised", e)
            ok = False
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
