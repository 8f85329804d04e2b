"""Per-CTA begin/end times of the lane decoder's two big kernels (ET_TUNE_DEBUG), one decode of text-1G
(developer tool, run under gpurun)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import entreepy_b200 as et  # noqa: E402
from entreepy_b200 import _abi, synth  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 1 << lg
man = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
thr = synth.thresholds_from_weights(synth.text_weights(man["midsummer_histogram"]))
c = et.Codec(0)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
c.synth_dev(dev.data_ptr(), n, synth.SEED, 0, thr)
enc = torch.empty(n + n // 8 + 16384, dtype=torch.uint8, device="cuda")
dec = torch.zeros(n, dtype=torch.uint8, device="cuda")
size = c.encode_dev(dev.data_ptr(), n, enc.data_ptr(), enc.numel())
for _ in range(3):
    c.decode_dev(enc.data_ptr() + 4, size - 4, dec.data_ptr(), n)
c.set_tuning(_abi.TUNE_DEBUG, 1)
c.decode_dev(enc.data_ptr() + 4, size - 4, dec.data_ptr(), n)
c.close()
