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


def _golden():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "driver_golden.json")))


def _gen_args(args):
    """the generator part of a golden command line (sizes, seeds, selectivity / skew, variant flags)"""
    toks, out, i = args.split(), [], 0
    while i < len(toks):
        if toks[i] in ("-r", "-s", "-q", "-x", "-y", "-z"):
            out += toks[i:i + 2]
            i += 2
        elif toks[i] in ("--non-unique", "--full-range"):
            out.append(toks[i])
            i += 1
        else:
            i += 1 if toks[i].startswith("--") else 2
    return out


@pytest.mark.parametrize("case", [c for c in _golden() if "--" in c["args"]], ids=lambda c: c["args"].replace(" ", ""))
def test_host_generators_reproduce_the_reference_binary(exe, oracle_mod, tmp_path, case):
    """--non-unique / --full-range / -z restate the reference's serial rand() generators (generator.c:531-651,
    genzipf.c:97-158): with the same seeds the driver must produce arrays whose join has exactly the `Results` and
    `S-tuples after filter` the UNMODIFIED reference binary printed (tests/golden/make_driver_golden.py). No GPU needed:
    the inputs are dumped in the -R/-S text format and joined by the oracle."""
    import numpy as np
    prefix = str(tmp_path / "rel_")
    p = run(exe, _gen_args(case["args"]) + ["--dump-relations", prefix])
    assert p.returncode == 0, p.stdout + p.stderr
    rel = {}
    for name in ("R", "S"):
        a = np.loadtxt(prefix + name + ".tbl", dtype=np.int64, skiprows=1).reshape(-1, 2)
        t = np.zeros(a.shape[0], dtype=oracle_mod.TUPLE)
        t["key"], t["payload"] = a[:, 0], a[:, 1]
        rel[name] = t
    toks = case["args"].split()
    opt = {toks[i]: toks[i + 1] for i in range(len(toks) - 1) if toks[i] in ("-b", "-m", "-k", "-B")}
    bloom = opt.get("-b", "no") != "no"
    res = oracle_mod.join(rel["R"], rel["S"], bloom, 1 if opt.get("-b") == "blocked" else 0, int(opt.get("-m", 1 << 28)),
                          int(opt.get("-k", 8)), int(opt.get("-B", 1024)))
    assert res["matches"] == case["results"]
    if case["filtered"] is not None:
        assert res["filtered"] == case["filtered"]
