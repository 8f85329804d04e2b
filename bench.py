#!/usr/bin/env python
"""Headline benchmark: .et encode + decode throughput (uncompressed GB/s) against the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A step = one encode (histogram -> host codebook -> pack) followed by one decode (dictionary parse
-> self-synchronising decode) of the workload, both through the C ABI (include/entreepy_b200.h).
`value` = uncompressed bytes per step / device time per step with the input resident in HBM;
`e2e` = the same through et_encode/et_decode with pinned HOST buffers (H2D and D2H inside the
timed region).  Workloads are BASELINE.md §4's (synthetic, splitmix64 seed 0xE7C0DE):
    N=1 default  text-1G  (2^30 B, the config the per-B200 roofline is quoted on)
    N>1 default  text-4G  (2^32-16 B, one .et stream sharded over the ranks: strong scaling)
Inputs are far larger than the 126 MB L2, so no explicit flush between iterations.

--impl reference times the reference ALGORITHM on the host (oracle/entreepy_oracle.c, a C
restatement: the reference is Zig and no Zig toolchain exists here or on the GPU box — there is no
oracle/_ref), single thread like the reference, on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "encode+decode round-trip GB/s (uncompressed)"
WORKLOADS = {  # name -> (bytes, weights kind)
    "midsummer": (112541, "file"),
    "text-5M": (5452595, "text"),
    "text-256M": (1 << 28, "text"),
    "text-1G": (1 << 30, "text"),
    "text-4G": ((1 << 32) - 16, "text"),
    "uniform256": (1 << 28, "uniform256"),
    "uniform255": (1 << 28, "uniform255"),
    "fib32": (1 << 28, "fib32"),
}
CPU_SAMPLE = 96 << 20       # bytes of the workload the cpu_baseline leg runs (about 10 s of host work)
REF_SAMPLE = 32 << 20       # bytes per step of the --impl reference arm (about 2.5 s per step)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def manifest():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))


def thresholds(kind):
    from entreepy_b200 import synth

    if kind == "text":
        w = synth.text_weights(manifest()["midsummer_histogram"])
    elif kind == "uniform256":
        w = synth.uniform_weights(0)
    elif kind == "uniform255":
        w = synth.uniform_weights(1)
    elif kind == "fib32":
        w = synth.fibonacci_weights(32)
    else:
        raise ValueError(kind)
    return synth.thresholds_from_weights(w)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_for(kernel):
    """dram bytes per launch of `kernel` from the committed ncu --set full capture, if any."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(kernel)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8:
                self.rows.append(f)

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.t.join(timeout=2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------- host baseline
def host_sample(kind, n):
    from entreepy_b200 import synth

    if kind == "file":
        return np.frombuffer(open(os.path.join(ROOT, "tests", "golden", "a_midsummer_nights_dream.txt"), "rb").read(),
                             dtype=np.uint8)[:n]
    return synth.generate(n, thresholds(kind))


def time_oracle(sample):
    """One encode + one decode of `sample` with the C restatement of the reference algorithm."""
    from oracle import oracle

    t0 = time.perf_counter()
    enc = oracle.encode(sample, cap=9000 + 5 * sample.size)
    t1 = time.perf_counter()
    rc, dec = oracle.decode_ref(enc[4:], sample.size)  # decode.zig's algorithm (hash probe per length)
    which = "decode.zig restatement"
    if rc != 0 or dec.size != sample.size:
        # the reference decoder cannot decode this stream: time the plain trie decoder instead
        t1 = time.perf_counter()
        dec = oracle.decode(enc[4:], sample.size)
        which = "bit-serial trie decoder (reference decoder fails on this input)"
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1, which


def run_reference(args, rank):
    if rank != 0:
        return 0
    n_total, kind = WORKLOADS[args.workload]
    n = min(n_total, REF_SAMPLE)
    sample = host_sample(kind, n)
    times = []
    which = ""
    for i in range(args.warmup + args.steps):
        te, td, which = time_oracle(sample)
        if i >= args.warmup:
            times.append((te, td))
    te = sum(t[0] for t in times) / len(times)
    td = sum(t[1] for t in times) / len(times)
    value = n / 1e9 / (te + td)
    sample_desc = f"first {n} B of {args.workload}, encode + decode per step; decode = {which}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (te + td) * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "bytes": n_total, "sample_bytes": n},
        "encode_gbs": n / 1e9 / te, "decode_gbs": n / 1e9 / td,
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample_desc,
                         "note": "C restatement of the reference algorithm (oracle/entreepy_oracle.c); the reference is "
                                 "Zig and cannot be built here (no Zig toolchain); its compute is single-threaded"},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world):
    import torch

    import entreepy_b200 as et
    from entreepy_b200 import _abi, sharded, synth

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the host baseline)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    codec = et.Codec(local)
    n_total, kind = WORKLOADS[args.workload]
    plan = sharded.ShardPlan(n_total, world, rank)
    n = plan.n_local
    stream = torch.cuda.current_stream().cuda_stream

    # ---- synthetic input, generated on the device (same bytes as entreepy_b200/synth.py on the CPU)
    inp = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    if kind == "file":
        data = host_sample(kind, n_total)[plan.lo:plan.hi]
        inp[:n].copy_(torch.from_numpy(data.copy()))
    else:
        codec.synth_dev(inp.data_ptr(), n, synth.SEED, plan.lo, thresholds(kind))
    enc = torch.empty(n + 16384, dtype=torch.uint8, device="cuda")
    dec = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    enc_flags = _abi.FLAG_WRITE_OUTPUT | _abi.FLAG_TIMING
    dec_flags = _abi.FLAG_WRITE_OUTPUT | _abi.FLAG_TIMING
    coder = sharded.ShardedCodec(codec, plan, dist)

    def step(stats=None):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        res = coder.encode(inp.data_ptr(), enc.data_ptr(), enc.numel(), enc_flags, stream)
        ms_enc = codec.last_stage_ms()
        e1.record()
        got = coder.decode(res, enc.data_ptr(), dec.data_ptr(), n, dec_flags, stream)
        ms_dec = codec.last_stage_ms() + [codec.last_decode_rounds]
        e2.record()
        if stats is not None:
            stats.append((e0, e1, e2, ms_enc, ms_dec))
        return res, got

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        res, got = step()
    barrier()
    assert got == n, f"decode produced {got} of {n} bytes"
    verified = bool(torch.equal(dec[:n], inp[:n]))  # round trip == original (north_star)
    if not verified:
        raise SystemExit("bench.py: decode(encode(x)) != x — refusing to report a throughput")

    launches0 = codec.kernel_launches
    stats = []
    with ClockSampler(local) as clocks:
        barrier()
        t_begin = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_begin.record()
        for _ in range(args.steps):
            step(stats)
        t_end.record()
        barrier()
    total_ms = t_begin.elapsed_time(t_end)
    launches = codec.kernel_launches - launches0
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_step = total_ms / args.steps
    enc_ms = [s[0].elapsed_time(s[1]) for s in stats]
    dec_ms = [s[1].elapsed_time(s[2]) for s in stats]
    hist_ms = statistics.mean(s[3][0] for s in stats)
    host_ms = statistics.mean(s[3][1] for s in stats)
    pack_ms = statistics.mean(s[3][2] for s in stats)
    unpack_ms = statistics.mean(s[4][2] for s in stats)
    c_local = res.local_bytes  # compressed bytes this rank wrote
    peak, peak_src = peaks()

    def roof(kernel, alg_bytes, ms):
        ach = alg_bytes / 1e9 / (ms / 1e3) if ms > 0 else 0.0
        return {"kernel": kernel, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic_for(kernel), "algorithmic_bytes": alg_bytes, "ms": ms, "peak_source": peak_src}

    roofs = [roof("histogram_kernel", n, hist_ms), roof("pack_kernel", n + c_local, pack_ms),
             roof("unpack_kernel", c_local + n, unpack_ms)]
    dominant = max(roofs, key=lambda r: r["ms"])

    # ---- end to end through the host-buffer entry points (pinned memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, codec, coder, inp, n, res, dist, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        k = min(n, CPU_SAMPLE)
        sample = inp[:k].cpu().numpy()
        te, td, which = time_oracle(sample)
        cpu = {"value": k / 1e9 / (te + td), "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": f"first {k} B of {args.workload}, one encode + one decode; decode = {which}",
               "encode_gbs": k / 1e9 / te, "decode_gbs": k / 1e9 / td,
               "note": "C restatement of the reference algorithm; the Zig reference cannot be built here"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": n_total / 1e9 / (ms_step / 1e3), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": args.workload, "bytes": n_total, "bytes_per_rank": n, "compressed_bytes_rank0": c_local,
                       "sharding": f"{world} contiguous byte ranges of one .et stream" if world > 1 else "none",
                       "l2": "inputs larger than L2 (126 MB), no explicit flush"},
            "encode_gbs": n_total / 1e9 / (statistics.mean(enc_ms) / 1e3),
            "decode_gbs": n_total / 1e9 / (statistics.mean(dec_ms) / 1e3),
            "encode_ms": statistics.mean(enc_ms), "decode_ms": statistics.mean(dec_ms),
            "stage_ms": {"histogram": hist_ms, "host_codebook": host_ms, "pack": pack_ms, "unpack": unpack_ms},
            "roofline": dominant, "rooflines": roofs,
            "encode_frac_of_hbm": (2 * n + c_local) / 1e9 / (statistics.mean(enc_ms) / 1e3) / peak,
            "decode_frac_of_hbm": (n + c_local) / 1e9 / (statistics.mean(dec_ms) / 1e3) / peak,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "verified_round_trip": verified,
            "decode_check_rounds": stats[-1][4][4],
            "clocks": clocks.summary(),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    codec.close()
    return 0


def run_e2e(args, codec, coder, inp, n, res, dist, barrier):
    """Same step through et_encode / et_decode with pinned host buffers."""
    import torch

    import entreepy_b200 as et

    h_in = codec.pinned(n)
    h_enc = codec.pinned(n + 16384)
    h_dec = codec.pinned(n)
    torch.from_numpy(h_in)[:] = inp[:n].cpu()
    ef = et.EncodeFlags(write_output=True)
    df = et.DecodeFlags(write_output=True)
    steps = max(1, min(args.steps, 3))
    size = 0
    times = []
    for i in range(1 + steps):
        barrier()
        t0 = time.perf_counter()
        size, info = coder.encode_host(h_in, h_enc, ef)
        got = coder.decode_host(info, h_enc, h_dec, df)
        barrier()
        if i >= 1:
            times.append(time.perf_counter() - t0)
    assert got == n and np.array_equal(h_dec[: 1 << 20], h_in[: 1 << 20])
    sec = statistics.mean(times)
    if dist is not None:
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    n_total = coder.plan.n_total
    # per step: encode uploads n and reads back `size`; decode uploads `size` (minus magic) and reads back n
    return {"value": n_total / 1e9 / sec, "unit": "GB/s", "h2d_bytes_per_step": int(n + size - 4),
            "d2h_bytes_per_step": int(size + n), "steps": steps, "ms_per_step": sec * 1e3,
            "path": "et_encode + et_decode (C ABI, pinned host buffers), wall clock around the blocking calls"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        args.workload = "text-1G" if max(world, args.gpus) == 1 else "text-4G"
    if args.impl == "reference":
        return run_reference(args, rank)
    if world != args.gpus:
        log(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE (launch N>1 under torch.distributed.run)")
    return run_ours(args, rank, world)


if __name__ == "__main__":
    sys.exit(main())
