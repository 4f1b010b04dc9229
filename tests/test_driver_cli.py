"""The C host driver's command line without a GPU: the reference's knobs (main.c:370-393,568-712) are parsed and
validated on the host, and a join without a device fails loudly instead of falling back to the CPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(H):
    from hwbloomradixjoin_b200 import build
    path = build.build_driver()
    assert path and os.path.exists(path)
    return path


def run(exe, args):
    return subprocess.run([exe] + args, capture_output=True, text=True, timeout=120)


def test_usage_lists_the_reference_knobs(exe):
    p = run(exe, ["-h"])
    assert p.returncode == 0
    for knob in ("-a", "-n", "-r", "-s", "-x", "-y", "-q", "-z", "-R", "-S", "-b", "-m", "-k", "-B"):
        assert f"  {knob} " in p.stdout or f" {knob} --" in p.stdout, knob
    for default in ("[PRO]", "[128000000]", "[12345]", "[54321]", "[1.0]", "[0.0]"):  # main.c:370-393
        assert default in p.stdout, default


def test_filter_arguments_are_checked_like_the_reference(exe):
    p = run(exe, ["-r", "1000", "-s", "1000", "-b", "basic", "-m", "1000"])
    assert p.returncode != 0 and "m must be a power of 2" in p.stdout          # bloom_filter.c:27-28
    p = run(exe, ["-r", "1000", "-s", "1000", "-b", "blocked", "-m", "1024", "-B", "48"])
    assert p.returncode != 0 and "B must be a power 2" in p.stdout             # bloom_filter.c:30-31
    p = run(exe, ["-r", "1000", "-s", "1000", "-b", "blocked", "-m", "1024", "-B", "2048"])
    assert p.returncode != 0 and "m must be a multiple of B" in p.stdout       # bloom_filter.c:32-33


def test_unknown_algorithm_is_reported(exe):
    p = run(exe, ["-a", "NOPE"])
    assert "does not exist" in p.stdout                                         # main.c:629-633


def test_join_without_a_gpu_fails_loudly(exe, H):
    if H.device_count() >= 1:
        pytest.skip("a GPU is present: covered by tests/test_gpu_driver.py")
    p = run(exe, ["-r", "1000", "-s", "1000"])
    assert p.returncode != 0
    assert "no CUDA device" in p.stdout and "no CPU fallback" in p.stdout
