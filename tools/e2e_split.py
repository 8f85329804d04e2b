"""Times et_encode and et_decode separately with pinned host buffers (developer tool, run under gpurun)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import entreepy_b200 as et
from entreepy_b200 import synth
man = json.load(open(os.path.join(ROOT, "tests/golden/manifest.json")))
thr = synth.thresholds_from_weights(synth.text_weights(man["midsummer_histogram"]))
n = 1 << 30
codec = et.Codec(0)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
codec.synth_dev(dev.data_ptr(), n, synth.SEED, 0, thr)
h_in, h_et, h_out = codec.pinned(n), codec.pinned(n + 16384), codec.pinned(n)
torch.from_numpy(h_in)[:] = dev.cpu()
for it in range(3):
    t0 = time.perf_counter(); size = codec.encode_into(h_in, h_et); t1 = time.perf_counter()
    got = codec.decode_into(h_et[4:size], h_out); t2 = time.perf_counter()
    print(f"encode {1e3*(t1-t0):.1f} ms  decode {1e3*(t2-t1):.1f} ms  size {size} got {got}")
# raw copies for reference
d = torch.empty(n, dtype=torch.uint8, device="cuda")
hp = torch.from_numpy(h_in); ho = torch.from_numpy(h_out)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(hp, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    ho.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"H2D 1GiB {1e3*(t1-t0):.1f} ms ({n/1e9/(t1-t0):.1f} GB/s)  D2H {1e3*(t2-t1):.1f} ms ({n/1e9/(t2-t1):.1f} GB/s)")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
with torch.cuda.stream(s1): d.copy_(hp, non_blocking=True)
with torch.cuda.stream(s2): ho.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"H2D + D2H concurrently, 1 GiB each: {1e3*(t1-t0):.1f} ms")
