"""The entreepy command-line driver (grammar of src/main.zig:42-208) over the C ABI."""
import os
import shutil
import subprocess

import pytest

from conftest import FIXTURES, GOLDEN, ROOT

CLI = os.path.join(ROOT, "entreepy_b200", "bin", "entreepy")


def run(*args, cwd=None):
    return subprocess.run([CLI, *args], capture_output=True, cwd=cwd, timeout=120)


@pytest.fixture(scope="module", autouse=True)
def built():
    from entreepy_b200 import build

    build.build()
    assert os.path.exists(CLI)


def test_help_and_argument_errors():
    assert b"Usage: entreepy" in run().stdout                       # no arguments -> help (main.zig:148-152)
    assert b"-o, --output" in run("-h").stdout
    assert b"Usage: entreepy" in run("--help").stdout
    r = run("-x", "c", "f")
    assert r.returncode == 1 and b"InvalidOption" in r.stderr       # main.zig:117
    r = run("frobnicate")
    assert r.returncode == 1 and b"InvalidCommand" in r.stderr      # main.zig:132
    r = run("--bogus")
    assert r.returncode == 1 and b"InvalidOption" in r.stderr       # main.zig:112
    r = run("c", "/nonexistent/file")
    assert r.returncode == 1 and b"FileNotFound" in r.stderr


@pytest.mark.gpu
def test_compress_and_decompress_files(tmp_path):
    for name in FIXTURES:
        src = tmp_path / name
        shutil.copy(os.path.join(GOLDEN, name), src)
        # options may come before or after the command; default output is [file].et
        r = run("c", str(src))
        assert r.returncode == 0, r.stderr
        et_path = str(src) + ".et"
        assert open(et_path, "rb").read() == open(os.path.join(GOLDEN, name + ".et"), "rb").read()
        assert b"=>" in r.stderr                                   # "X => Y" summary on stderr (encode.zig:334)
        # default decode output: decoded_[file] next to the input (main.zig:160-169, fixed)
        r = run("d", et_path)
        assert r.returncode == 0, r.stderr
        assert open(tmp_path / ("decoded_" + name), "rb").read() == open(src, "rb").read()
        # explicit output, combined flags, print to stdout
        out = tmp_path / "explicit.txt"
        r = run("-pd", "d", et_path, "-o", str(out))
        assert r.returncode == 0, r.stderr
        assert open(out, "rb").read() == open(src, "rb").read()
        assert open(src, "rb").read() in r.stdout                   # -p streams the text to stdout (decode.zig:189)
        assert b"time taken:" in r.stdout                           # -d (decode.zig:16)


@pytest.mark.gpu
def test_dry_run_writes_no_file_and_debug_prints_dictionary(tmp_path):
    src = tmp_path / "test.txt"
    shutil.copy(os.path.join(GOLDEN, "test.txt"), src)
    r = run("-td", "c", str(src), "--output", str(tmp_path / "never.et"))
    assert r.returncode == 0, r.stderr
    assert not os.path.exists(tmp_path / "never.et")               # -t: main.zig:191
    assert b"bits in output: 336" in r.stdout                       # 42 bytes (encode.zig:320)
    assert b"D 68 - 00" in r.stdout and b"C 67 - 11101" in r.stdout  # dictionary dump (encode.zig:205-211)


@pytest.mark.gpu
def test_strict_rejects_wrong_magic_and_empty_input_is_queue_empty(tmp_path):
    bad = tmp_path / "bad.et"
    good = open(os.path.join(GOLDEN, "test.txt.et"), "rb").read()
    bad.write_bytes(b"\x00" + good[1:])
    assert run("--strict", "-t", "d", str(bad)).returncode == 1
    assert run("-t", "d", str(bad)).returncode == 0                 # the reference never looks at file[0..4)
    empty = tmp_path / "empty.txt"
    empty.write_bytes(b"")
    r = run("-t", "c", str(empty))
    assert r.returncode == 1 and b"QueueEmpty" in r.stderr          # encode.zig:138
