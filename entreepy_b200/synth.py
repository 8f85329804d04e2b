"""Synthetic inputs of BASELINE.md §4 (CPU twin of et_synth_dev, identical bytes).

byte[i] = smallest s with thresholds[s] > (splitmix64(seed + i) >> 32), thresholds being the
cumulative distribution scaled to 2^32.  One PRNG, one seed (0xE7C0DE), CPU and GPU agree.
"""
import numpy as np

SEED = 0xE7C0DE
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return x ^ (x >> np.uint64(31))


def thresholds_from_weights(weights):
    """weights[256] (any non-negative ints) -> uint32[256] cumulative thresholds."""
    w = np.asarray(weights, dtype=object)
    total = int(sum(int(v) for v in w))
    assert total > 0
    thr = np.zeros(256, dtype=np.uint32)
    acc = 0
    for s in range(256):
        acc += int(w[s])
        thr[s] = min((acc << 32) // total, 0xFFFFFFFF)
    # every value r < 2^32 must land on a symbol with non-zero weight
    last = max(s for s in range(256) if int(w[s]) > 0)
    thr[last:] = 0xFFFFFFFF
    return thr


def generate(n, thresholds, seed=SEED, first_index=0):
    """numpy uint8[n]; chunked so memory stays bounded."""
    out = np.empty(n, dtype=np.uint8)
    thr = np.asarray(thresholds, dtype=np.uint32).astype(np.uint64)
    step = 1 << 22
    with np.errstate(over="ignore"):
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            idx = np.arange(lo, hi, dtype=np.uint64) + np.uint64((seed + first_index) & 0xFFFFFFFFFFFFFFFF)
            r = splitmix64(idx) >> np.uint64(32)
            out[lo:hi] = np.minimum(np.searchsorted(thr, r, side="right"), 255).astype(np.uint8)
    return out


def text_weights(midsummer_histogram):
    """English-frequency text: i.i.d. bytes from the 93-symbol histogram of A Midsummer Night's Dream."""
    return list(midsummer_histogram)


def uniform_weights(first=0):
    """Uniform over byte values first..255 (first=0: all 256, first=1: the 255-symbol variant)."""
    return [0] * first + [1] * (256 - first)


def fibonacci_weights(depth=32):
    """depth+1 symbols with weights F(1)..F(depth+1): the Huffman tree is a chain of that depth."""
    f = [1, 1]
    while len(f) < depth + 1:
        f.append(f[-1] + f[-2])
    return f[: depth + 1] + [0] * (256 - depth - 1)


def fibonacci_counts(n, depth=32):
    """Exact symbol counts for the maximum-depth configuration: depth+1 symbols whose counts are
    F(1)..F(depth+1) times the largest factor that fits n, the remainder added to the most frequent one.
    Scaling keeps the tree a chain, so the longest code has `depth` bits — i.i.d. sampling from the same
    weights does not (the rarest symbols come out with the wrong ratios, or not at all)."""
    f = [1, 1]
    while len(f) < depth + 1:
        f.append(f[-1] + f[-2])
    f = f[: depth + 1]
    k = n // sum(f)
    if k < 1:
        raise ValueError("n too small for a chain of this depth")
    counts = [k * x for x in f] + [0] * (256 - depth - 1)
    counts[depth] += n - k * sum(f)
    return counts


def shuffled_dev(counts, seed=SEED, device="cuda"):
    """uint8 CUDA tensor with exactly counts[s] copies of every symbol s, in a seeded random order (torch's
    generator: reproducible on the same device type, not the CPU twin of anything)."""
    import torch

    c = torch.tensor(list(counts), dtype=torch.int64, device=device)
    runs = torch.repeat_interleave(torch.arange(256, dtype=torch.uint8, device=device), c)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) & 0x7FFFFFFF)
    perm = torch.randperm(runs.numel(), generator=g, device=device)
    out = runs[perm]
    torch.cuda.synchronize()  # callers hand the pointer to the codec, which works on its own stream
    return out
