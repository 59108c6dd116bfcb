"""The streaming host path's flow control (dbeel_b200/csrc/host/stream_pump.h) on a box without a GPU: a fake engine
loop + in-memory files, every ring size / thread count, error injection (tests/stream_pump_test.cc)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_stream_pump_moves_every_byte_once_and_stops_on_errors(tmp_path):
    exe = str(tmp_path / "stream_pump_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(HERE, "stream_pump_test.cc"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("ok")
