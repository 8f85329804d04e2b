"""Builds tuning variants of the library (different table windows / split) and, on the GPU box, times each with
tools/dec_bench.py.   python tools/variants.py build | run [log2_bytes] [kind]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    "c15w14": ["ET_COUNT_BITS=15", "ET_WRITE_BITS=14"],
    "l25": ["ET_LANE_WORDS=25"],
    "l29": ["ET_LANE_WORDS=29"],
    "l21": ["ET_LANE_WORDS=21"],
}
if sys.argv[1] == "build":
    from entreepy_b200 import build

    for tag, defs in VARIANTS.items():
        print(build.build_variant(tag, defs))
else:
    lg = sys.argv[2] if len(sys.argv) > 2 else "30"
    kind = sys.argv[3] if len(sys.argv) > 3 else "text"
    for tag in VARIANTS:
        env = dict(os.environ, ET_LIB=os.path.join(ROOT, "entreepy_b200", "lib", f"libentreepy_b200_{tag}.so"))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dec_bench.py"), lg, "7", kind], env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(tag, "decode_ms", round(d["decode_ms"], 4), "min", round(d["decode_min_ms"], 4), "ok", d["round_trip_ok"], flush=True)
        except Exception:
            print(tag, "FAILED", out.stdout[-300:], out.stderr[-600:], flush=True)
