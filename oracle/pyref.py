"""Second, independent restatement of the reference encoder in pure Python (small inputs only).

TEST INFRASTRUCTURE.  Written separately from entreepy_oracle.c so that agreement between
the two is a cross-check of the source reading (encode.zig:43-318, queue.zig:9-43).
"""
from collections import deque


def dictionary(counts):
    """counts[256] -> {sym: (data_u32, length)} for symbols that became leaves (encode.zig:54-214)."""
    order = sorted((s for s in range(256) if counts[s] > 0), key=lambda s: (counts[s], s))
    if len(order) == 256:  # encode.zig:70/79: u8 index saturates, used as exclusive bound
        order = order[:255]
    leaves = deque((counts[s], ("leaf", s)) for s in order)
    saplings = deque()

    def take():
        if not saplings:
            return leaves.popleft()
        if not leaves:
            return saplings.popleft()
        return leaves.popleft() if leaves[0][0] <= saplings[0][0] else saplings.popleft()  # encode.zig:113

    while len(leaves) + len(saplings) > 1:
        a = take()
        b = take()
        saplings.append((a[0] + b[0], ("node", a[1], b[1])))  # first pick = left (encode.zig:124)
    if not leaves and not saplings:
        raise IndexError("QueueEmpty")  # encode.zig:138
    root = (leaves or saplings)[0][1]
    out = {}
    todo = [(root, 0, 0)]
    while todo:
        node, data, length = todo.pop()
        if node[0] == "leaf":
            out[node[1]] = (data & 0xFFFFFFFF, length)
        else:
            todo.append((node[2], ((data << 1) | 1) & 0xFFFFFFFF, length + 1))  # right = 1 (encode.zig:181-183)
            todo.append((node[1], (data << 1) & 0xFFFFFFFF, length + 1))  # left = 0 (encode.zig:195-197)
    return out


def _emitted_bits(data, length):
    return [(data >> ((j - 1) & 31)) & 1 for j in range(length, 0, -1)]  # encode.zig:293,311


def encode(text: bytes) -> bytes:
    counts = [0] * 256
    for c in text:
        counts[c] += 1
    d = dictionary(counts)
    bits = []

    def put(v, n):
        bits.extend((v >> (n - 1 - i)) & 1 for i in range(n))

    put(0xE7C0DE, 24)
    put(1, 8)
    live = [s for s in range(256) if s in d and d[s][1] > 0]
    put(max(len(live) - 1, 0), 8)
    put(len(text) & 0xFFFFFFFF, 32)
    for s in live:
        put(s, 8)
        put(d[s][1], 8)
        bits.extend(_emitted_bits(*d[s]))
    bits.extend([0] * (-len(bits) % 8))
    for c in text:
        if c in d:
            bits.extend(_emitted_bits(*d[c]))
    bits.extend([0] * (-len(bits) % 8))
    out = bytearray(len(bits) // 8)
    for i, b in enumerate(bits):
        if b:
            out[i >> 3] |= 0x80 >> (i & 7)
    return bytes(out)
