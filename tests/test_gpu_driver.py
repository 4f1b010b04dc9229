"""The C host driver (host/mchashjoins_gpu.c) on a GPU: same stdout contract as the reference's mchashjoins, i.e. the
regular expressions of measurements/run.py:109-129 must parse it, and the numbers must be the golden ones."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(args):
    from hwbloomradixjoin_b200 import build
    build.build_library()
    exe = build.build_driver()
    p = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    return p.stdout


def test_driver_stdout_contract(Hgpu):
    out = run("-a PRO -n 8 -r 250000 -s 2000000 -q 0.01 -b basic -m 2097152 -k 1".split())
    # the lines run.py parses (measurements/run.py:109-129)
    assert re.search(r"relation S with size = [\d.]+ MiB, #tuples = 2000000 : OK", out)
    assert re.search(r"S-tuples after filter: 241986", out)
    m = re.search(r"RUNTIME TOTAL, BUILD, PART \(cycles\): \n(\d+) \t (\d+) \t (\d+)", out)
    assert m
    m = re.search(r"TOTAL-TIME-USECS, TOTAL-TUPLES, NSEC-PER-TUPLE: \n([\d.]+) \t (\d+) \t ([\d.]+)", out)
    assert m and int(m.group(2)) == 20000 and float(m.group(1)) > 0
    assert re.search(r"PARTITION-TIME-USECS, PROBE-TIME-USECS, JOIN-TIME-USECS: \n[\d.]+ \t [\d.]+\t [\d.]+", out)
    assert "[INFO ] Results = 20000. DONE." in out
    assert "[INFO ] Running join algorithm PRO ..." in out


def test_driver_variants(Hgpu):
    out = run("-a RJ -r 250000 -s 2000000 -q 0.01 -b blocked -m 2097152 -k 3 -B 512".split())
    assert "S-tuples after filter" not in out  # BRJ does not print it
    assert "Results = 20000." in out
    out = run("-a PRH -r 250000 -s 2000000 -q 0.01 -b blocked -m 2097152 -k 3 -B 512".split())
    assert "S-tuples after filter: 74851" in out
    out = run("-r 250000 -s 1000000".split())  # defaults: plain PRO, q = 1.0
    assert "Results = 1000000." in out and "S-tuples after filter" not in out
    out = run("-r 100000 -s 500000 -z 1.0 -b basic -m 1048576 -k 1".split())  # Zipf: everything matches and passes
    assert "S-tuples after filter: 500000" in out and "Results = 500000." in out


def test_driver_rejects_bad_arguments(Hgpu):
    from hwbloomradixjoin_b200 import build
    exe = build.build_driver()
    p = subprocess.run([exe, "-r", "1000", "-s", "1000", "-b", "basic", "-m", "1000"], capture_output=True, text=True)
    assert p.returncode != 0 and "m must be a power of 2" in p.stdout
    p = subprocess.run([exe, "-a", "NOPE"], capture_output=True, text=True)
    assert "does not exist" in p.stdout


def _golden():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "driver_golden.json")))


@pytest.mark.parametrize("case", _golden(), ids=lambda c: c["args"].replace(" ", ""))
def test_driver_matches_the_reference_binary_on_seeded_generators(Hgpu, case):
    """--non-unique, --full-range and -z inputs depend on glibc rand() and the seeds: the same command line must print
    what the UNMODIFIED reference binary printed (tests/golden/driver_golden.json)"""
    out = run(case["args"].split())
    assert f"Results = {case['results']}." in out
    if case["filtered"] is not None and " RJ " not in f" {case['args']} ":
        assert f"S-tuples after filter: {case['filtered']}" in out


def test_driver_loads_relations_from_files(Hgpu, tmp_path):
    """-R / -S (load_relation, generator.c:418-436,686-741): dump seeded inputs, load them back, same result"""
    prefix = str(tmp_path / "rel_")
    gen = "-r 120000 -s 600000 -q 0.3 --non-unique -x 5 -y 6".split()
    run(gen + ["--dump-relations", prefix])
    direct = run(gen + "-b basic -m 1048576 -k 1".split())
    loaded = run(["-r", "120000", "-s", "600000", "-R", prefix + "R.tbl", "-S", prefix + "S.tbl"] + "-b basic -m 1048576 -k 1".split())
    pick = lambda o: (re.search(r"Results = (\d+)", o).group(1), re.search(r"after filter: (\d+)", o).group(1))
    assert pick(direct) == pick(loaded)
    assert "Loading relation R" in loaded


def test_driver_on_several_gpus(Hgpu):
    if Hgpu.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    one = run("-a PRO -r 250000 -s 2000000 -q 0.01 -b basic -m 2097152 -k 1".split())
    two = run("-a PRO -r 250000 -s 2000000 -q 0.01 -b basic -m 2097152 -k 1 --gpus 2".split())
    assert "S-tuples after filter: 241986" in one and "S-tuples after filter: 241986" in two
    assert "Results = 20000." in two
    assert re.search(r"H2D-COPY-USECS, END-TO-END-USECS, GPUS: \n[\d.]+ \t [\d.]+\t 2 ", two)
