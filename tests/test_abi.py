"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/hwbrj.h
declares, the ctypes mirrors have the C layouts, and the host-side mirror of the reference interface behaves like
the reference's (argument validation, dispatch table). No compute call is made here (no GPU in this tier)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hwbrj.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"static inline.*?\n}\n", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_header_declares_reference_entry_points():
    names = _declared_functions()
    for n in ("BPRO", "BRJ", "BPRH", "BPRHO", "PRO", "RJ", "PRH", "PRHO"):
        assert n in names


def test_library_exports_every_declared_symbol(H):
    from hwbloomradixjoin_b200 import _native as N
    L = N.load()
    declared = _declared_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/hwbrj.h but not exported"
    assert sorted(N.EXPORTS) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", N.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf" T {name}\b", out), name


def test_header_compiles_as_plain_c(tmp_path):
    """the boundary is a C ABI: a C11 translation unit including the header must compile (no C++/torch types)"""
    src = tmp_path / "t.c"
    src.write_text('#include "hwbrj.h"\n'
                   '_Static_assert(sizeof(tuple_t) == 8, "tuple");\n'
                   '_Static_assert(sizeof(relation_t) == 16, "relation");\n'
                   '_Static_assert(sizeof(result_t) == 24, "result");\n'
                   '_Static_assert(sizeof(bloom_filter_args_t) == 32, "args");\n'
                   'int main(void){ relation_t r = {0,0}; (void)r; return (int)hwbrj_mix64(1,2) & 0; }\n')
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                        str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_ctypes_layouts_match_c(H, tmp_path):
    from hwbloomradixjoin_b200 import _native as N
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "hwbrj.h"\n'
                   'int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(tuple_t), sizeof(relation_t),'
                   ' sizeof(result_t), sizeof(bloom_filter_args_t), sizeof(hwbrj_stats_t), offsetof(hwbrj_stats_t, ms_total),'
                   ' offsetof(hwbrj_stats_t, h2d_bytes), offsetof(hwbrj_stats_t, kernel_launches)); return 0; }\n')
    exe = tmp_path / "sz"
    r = subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sizes = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()))
    assert sizes == [C.sizeof(N.TupleT), C.sizeof(N.RelationT), C.sizeof(N.ResultT), C.sizeof(N.BloomFilterArgsT),
                     C.sizeof(N.StatsT), N.StatsT.ms_total.offset, N.StatsT.h2d_bytes.offset,
                     N.StatsT.kernel_launches.offset]


def test_mix64_python_matches_header(H, tmp_path):
    src = tmp_path / "m.c"
    src.write_text('#include <stdio.h>\n#include "hwbrj.h"\nint main(void){ printf("%llu\\n",'
                   ' (unsigned long long)hwbrj_mix64(123456u, 4000000000u)); return 0; }\n')
    exe = tmp_path / "m"
    assert subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)]).returncode == 0
    got = int(subprocess.run([str(exe)], capture_output=True, text=True).stdout)
    z = ((123456 << 32) | 4000000000) + 0x9e3779b97f4a7c15 & (2**64 - 1)
    z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & (2**64 - 1)
    assert got == z ^ (z >> 31)


def test_check_args_like_reference(H):
    """assert_args (bloom_filter.c:26-34): m power of two; for BLOCKED, B power of two and m % B == 0."""
    from hwbloomradixjoin_b200 import _native as N
    L = N.load()

    def rc(variant, m, k, B):
        a = N.BloomFilterArgsT(variant, m, k, B)
        return L.hwbrj_check_args(C.byref(a))

    assert rc(0, 1 << 30, 1, 512) == 0
    assert rc(0, (1 << 30) + 8, 1, 512) != 0
    assert rc(0, 1 << 20, 1, 500) == 0          # B is ignored for BASIC
    assert rc(1, 1 << 20, 3, 500) != 0
    assert rc(1, 1 << 20, 3, 1 << 21) != 0      # m % B
    assert rc(1, 1 << 20, 3, 256) == 0
    assert rc(0, 1 << 33, 1, 512) != 0          # uint32 size arithmetic of the reference (bloom_filter.c:60-63)
    for bad in [H.BloomFilterArgs(0, 1000, 1, 512), H.BloomFilterArgs(1, 1 << 20, 1, 48), H.BloomFilterArgs(1, 1 << 10, 1, 1 << 11)]:
        with pytest.raises(ValueError):
            bad.check()
    H.BloomFilterArgs().check()  # CLI defaults main.c:389-393
    d = H.BloomFilterArgs()
    assert (d.variant, d.m, d.k, d.B) == (0, 256 << 20, 8, 1024)


def test_dispatch_table_mirrors_main_c(H):
    assert set(H.ALGOS) == {"PRO", "RJ", "PRH", "PRHO"}  # main.c:331-339 minus the out-of-scope NPO pair
    assert H.ALGOS["PRO"] == (H.PRO, H.BPRO) and H.ALGOS["RJ"] == (H.RJ, H.BRJ)
    with pytest.raises(KeyError):
        H.run("NOPE", np.zeros(1, H.TUPLE), np.zeros(1, H.TUPLE))


def test_relation_coercion(H):
    from hwbloomradixjoin_b200.api import as_relation
    a = np.arange(10, dtype=np.int32).reshape(5, 2)
    r = as_relation(a)
    assert r.dtype == H.TUPLE and r["key"].tolist() == [0, 2, 4, 6, 8] and r["payload"].tolist() == [1, 3, 5, 7, 9]
    with pytest.raises(TypeError):
        as_relation(np.zeros(4, dtype=np.float32))


def test_no_cpu_fallback_without_gpu(H):
    """Without a device the product path must fail loudly, not compute on the CPU."""
    if H.device_count() > 0:
        pytest.skip("a GPU is present")
    code = ("import numpy as np, hwbloomradixjoin_b200 as H\n"
            "R=np.zeros(4,H.TUPLE); S=np.zeros(4,H.TUPLE)\n"
            "print(H.PRO(R,S,1).totalresults)\n")
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode != 0
    assert "no CUDA device" in r.stdout + r.stderr


def test_product_never_imports_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "hwbloomradixjoin_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(base, f)).read()
                assert "import oracle" not in txt and "liboracle" not in txt and "libref" not in txt, f
