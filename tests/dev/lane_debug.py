"""Bring-up check of the lane-interleaved decoder with first-difference diagnostics (developer tool, run under gpurun)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("ET_LANE_MIN_BYTES", "0")
os.environ.setdefault("ET_DEBUG_LANES", "1")
import entreepy_b200 as et  # noqa: E402
from entreepy_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    man = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
    thr = synth.thresholds_from_weights(synth.text_weights(man["midsummer_histogram"]))
    host = synth.generate(1 << 20, thr)
    c = et.Codec(0)
    for n in [int(x) for x in sys.argv[1:]] or [30011, 100003, 1 << 20]:
        data = host[:n]
        stream = oracle.encode(data, cap=9000 + 5 * n).tobytes()[4:]
        print("n", n, "stream", len(stream), flush=True)
        m, out = c.decode(stream)
        ok = m == n and out.tobytes() == data.tobytes()
        print("  ->", m, "ok" if ok else "MISMATCH", "rounds", c.last_decode_rounds, flush=True)
        if not ok:
            k = min(m, n)
            d = np.nonzero(out[:k] != data[:k])[0]
            print("  ndiff", d.size, "first", d[:10], flush=True)
            if d.size:
                i = int(d[0])
                print("  got ", out[max(0, i - 8): i + 24].tobytes())
                print("  want", data[max(0, i - 8): i + 24].tobytes())
            break
    c.close()


main()
