"""Re-runs one case of tests/test_gpu_fuzz.py with diagnostics (developer tool, run under gpurun): python tests/dev/fuzz_case.py IT"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import entreepy_b200 as et  # noqa: E402
from entreepy_b200 import _abi  # noqa: E402
from oracle import oracle  # noqa: E402
from test_gpu_fuzz import _case  # noqa: E402

target = int(sys.argv[1])
rng = np.random.default_rng(20261018)
c = et.Codec(0)
c.set_tuning(_abi.TUNE_LANE_MIN_BYTES, 0)
for it in range(target + 1):
    data = _case(rng)
    if np.unique(data).size < 2:
        continue
    in_off, out_off, dec_off = (int(x) for x in rng.integers(0, 16, 3))
    if it < target:
        continue
    want = oracle.encode(data, cap=9000 + 5 * data.size).tobytes()
    d = et.parse_header(want[4:])
    print("case", it, "n", data.size, "symbols", np.unique(data).size, "lengths", d.min_length, d.max_length, "body_offset", d.body_offset,
          "offs", in_off, out_off, dec_off)
    text = oracle.decode(want[4:], data.size)
    d_et = torch.zeros(len(want) + 64, dtype=torch.uint8, device="cuda")
    d_et[out_off : out_off + len(want)] = torch.from_numpy(np.frombuffer(want, dtype=np.uint8).copy()).cuda()
    for dbg in (0, 1):
        c.set_tuning(_abi.TUNE_DEBUG, dbg)
        d_out = torch.zeros(data.size + 32, dtype=torch.uint8, device="cuda")
        m = c.decode_dev(d_et.data_ptr() + out_off + 4, len(want) - 4, d_out.data_ptr() + dec_off, data.size)
        got = d_out[dec_off : dec_off + m].cpu().numpy()
        bad = np.nonzero(got != text[:m])[0]
        print("decoded", m, "of", text.size, "rounds", c.last_decode_rounds, "mismatches", bad.size, "first", bad[:5], "last", bad[-3:] if bad.size else None)
        if bad.size:
            k = int(bad[0])
            print(" got ", got[max(k - 4, 0) : k + 12])
            print(" want", text[max(k - 4, 0) : k + 12])
