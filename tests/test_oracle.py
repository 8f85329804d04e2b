"""The CPU oracle against every known answer the reference offers for this path."""
import hashlib

import numpy as np
import pytest

from conftest import FIXTURES, make_cases
from oracle import oracle, pyref


def test_golden_sizes_and_hashes(fixtures, manifest):
    for name in FIXTURES:
        et = oracle.encode(fixtures[name]).tobytes()
        m = manifest[name]
        assert len(et) == m["et_bytes"]
        assert hashlib.sha256(et).hexdigest() == m["et_sha256"]


def test_readme_known_answer(fixtures):
    # README.md:51 — "Macbeth, Act V, Scene V | 477 bytes | 374 bytes"
    assert len(fixtures["nice.shakespeare.txt"]) == 477
    assert oracle.encode(fixtures["nice.shakespeare.txt"]).size == 374


def test_hand_traced_vector(fixtures):
    # tree traced by hand from encode.zig:102-135 for res/test.txt
    want = bytes.fromhex(
        "e7c0de01060000002f0a05e208148409e860bd44021140f2f812" "934243dc619f87efccfe8c37dd9fc700"
    )
    assert oracle.encode(fixtures["test.txt"]).tobytes() == want
    occ = oracle.histogram(fixtures["test.txt"])
    data, length = oracle.build_dictionary(occ)
    codes = {chr(s): format(int(data[s]), "b").zfill(int(length[s])) for s in range(256) if length[s]}
    assert codes == {"D": "00", "_": "01", "A": "10", "E": "110", "B": "1111", "\n": "11100", "C": "11101"}


def test_golden_files_match_committed(fixtures, golden_et):
    for name in FIXTURES:
        assert oracle.encode(fixtures[name]).tobytes() == golden_et[name]
        assert golden_et[name][:4] == b"\xe7\xc0\xde\x01"  # README.md:59-62


def test_round_trip_fixtures(fixtures, golden_et):
    # test.zig:35-72: decode(encode(text)[4..]) == text, for both decoders
    for name in FIXTURES:
        text = fixtures[name]
        assert oracle.decode(golden_et[name][4:], len(text) * 2).tobytes() == text
        rc, out = oracle.decode_ref(golden_et[name][4:], len(text) * 2)
        assert rc == 0 and out.tobytes() == text


def test_two_restatements_agree():
    for name, data in make_cases().items():
        if data.size > 20000:
            data = data[:20000]
        assert pyref.encode(data.tobytes()) == oracle.encode(data).tobytes(), name


def test_empty_input_is_queue_empty():
    with pytest.raises(oracle.OracleError) as e:  # encode.zig:138
        oracle.encode(b"")
    assert e.value.code == oracle.ERR_QUEUE_EMPTY


def test_single_symbol_has_no_dictionary():
    et = oracle.encode(bytes([7]) * 1000).tobytes()
    assert et == b"\xe7\xc0\xde\x01\x00" + (1000).to_bytes(4, "big")  # 9 bytes, zero entries, empty body


def test_256_symbols_drop_the_last_in_sort_order():
    # encode.zig:70,79: u8 index saturates; the most frequent symbol (ties: highest byte) gets no code
    data = np.concatenate([np.arange(256, dtype=np.uint8), np.full(10, 200, np.uint8)])
    occ = oracle.histogram(data)
    _, n = oracle.sort_symbols(occ)
    assert n == 255
    _, length = oracle.build_dictionary(occ)
    assert length[200] == 0 and (length > 0).sum() == 255
    et = oracle.encode(data).tobytes()
    assert et[4] == 254
    out = oracle.decode(et[4:], data.size)  # lossy: the dropped symbol is gone from the stream
    assert out.size <= data.size and 200 not in set(out[: data.size - 11].tolist())


def test_code_data_truncates_past_32_bits():
    fib = [1, 1]
    while len(fib) < 36:
        fib.append(fib[-1] + fib[-2])
    occ = np.zeros(256, dtype=np.uint64)
    occ[: len(fib)] = fib
    data, length = oracle.build_dictionary(occ)
    assert int(length.max()) == 35 and all(int(d) < 2**32 for d in data)


def test_reference_decoder_defects_are_restated():
    # symbol 0x00 never matches (table value 0 = empty) and the loop spins
    data = np.array([0, 1, 1, 2, 0, 0, 1, 2, 2, 2] * 20, dtype=np.uint8)
    et = oracle.encode(data).tobytes()
    assert oracle.decode(et[4:], data.size).tobytes() == data.tobytes()
    rc, _ = oracle.decode_ref(et[4:], data.size)
    assert rc == oracle.ERR_HANG
