#!/usr/bin/env python
"""SASS evidence per kernel: disassembles libhwbrj_cuda.so (cuobjdump -sass) and writes, for every kernel of the join,
profiles/<prefix>_sass_<kernel>.txt with the instruction-class counts that show HOW the kernel moves data -- TMA bulk
copies (UBLKCP), mbarrier traffic (SYNCS), 128-bit streaming loads (LDG.E.128 / .NA / .CONSTANT), global reductions and
atomics (REDG / ATOMG, and whether they are system-scope), shared-memory atomics (ATOMS), votes and shuffles -- plus the
register / shared-memory footprint. A one-line-per-kernel summary goes to profiles/<prefix>_sass_summary.txt.

    python tools/sass_report.py [prefix=r2]        (runs here: no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hwbloomradixjoin_b200", "libhwbrj_cuda.so")

CLASSES = [
    ("UBLKCP", r"^UBLKCP"), ("UTMALDG", r"^UTMALDG"), ("SYNCS (mbarrier)", r"^SYNCS"),
    ("LDG.E.128", r"^LDG\.E\.(\w+\.)*128"), ("LDG.E.64", r"^LDG\.E\.(\w+\.)*64"), ("LDG (other)", r"^LDG"),
    ("LDG ... .CONSTANT / .NA (non-coherent / no L1 allocate)", None),
    ("STG.E.128", r"^STG\.E\.(\w+\.)*128"), ("STG.E.64", r"^STG\.E\.(\w+\.)*64"), ("STG (other)", r"^STG"),
    ("REDG (global reduction, no return)", r"^REDG"), ("ATOMG (global atomic)", r"^ATOMG"),
    ("... of which .SYS scope", None),
    ("ATOMS (shared atomic)", r"^ATOMS"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDSM/LDGSTS", r"^(LDSM|LDGSTS)"),
    ("BAR", r"^BAR"), ("VOTE", r"^VOTE"), ("SHFL", r"^SHFL"), ("POPC", r"^POPC"), ("MATCH", r"^MATCH"),
    ("IMAD.WIDE / IMAD.HI (hash multiplies)", r"^IMAD\.(WIDE|HI)"), ("MEMBAR / FENCE", r"^(MEMBAR|FENCE)"),
    ("CCTL", r"^CCTL"), ("LDL/STL (local memory: spills or indexed params)", r"^(LDL|STL)"),
]


def demangle(name: str) -> str:
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = tuple(map(int, m.groups()))
    kernels = collections.OrderedDict()
    cur = None
    arch = None
    for line in sass.splitlines():
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur].append(m.group(1))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    summary = [f"cuobjdump -sass hwbloomradixjoin_b200/libhwbrj_cuda.so  (arch = {arch}); one file per kernel: {prefix}_sass_<kernel>.txt",
               f"{'kernel':<44} {'instr':>6} {'REG':>4} {'SMEM':>6} {'UBLKCP':>6} {'SYNCS':>5} {'LDG128':>6} {'REDG':>5} {'ATOMG':>5} {'.SYS':>4} {'ATOMS':>5} {'VOTE':>4}"]
    for mangled, ops in kernels.items():
        name = demangle(mangled)
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("hwbrj::", "")
        if not short.startswith("k_"):
            continue
        counts = collections.OrderedDict()
        rest = list(ops)
        for label, pat in CLASSES:
            if pat is None:
                continue
            hit = [o for o in rest if re.match(pat, o)]
            rest = [o for o in rest if not re.match(pat, o)]
            counts[label] = len(hit)
        counts["LDG ... .CONSTANT / .NA (non-coherent / no L1 allocate)"] = sum(1 for o in ops if o.startswith("LDG") and (".CONSTANT" in o or ".NA" in o))
        counts["... of which .SYS scope"] = sum(1 for o in ops if (o.startswith("ATOMG") or o.startswith("REDG") or o.startswith("ST") or o.startswith("LD")) and ".SYS" in o)
        reg, stack, shared = usage.get(mangled, (0, 0, 0))
        fname = re.sub(r"[^A-Za-z0-9_]+", "_", short).strip("_")
        mix = collections.Counter(o.split(".")[0] for o in ops).most_common(14)
        with open(os.path.join(ROOT, "profiles", f"{prefix}_sass_{fname}.txt"), "w") as f:
            f.write(f"{name}\narch {arch}; {len(ops)} SASS instructions; REG {reg}, STACK {stack} B, static SHARED {shared} B\n\n")
            for label, _ in CLASSES:
                f.write(f"  {label:<58} {counts[label]:>6}\n")
            f.write("\n  opcode mix (top 14): " + ", ".join(f"{k} {v}" for k, v in mix) + "\n")
            distinct = sorted(set(o for o in ops if re.match(r"^(LDG|STG|ATOMG|REDG|UBLKCP|SYNCS|ATOMS|LD\.|ST\.)", o)))
            f.write("  memory opcodes seen: " + ", ".join(distinct) + "\n")
        summary.append(f"{short:<44} {len(ops):>6} {reg:>4} {shared:>6} {counts['UBLKCP']:>6} {counts['SYNCS (mbarrier)']:>5} "
                       f"{counts['LDG.E.128']:>6} {counts['REDG (global reduction, no return)']:>5} {counts['ATOMG (global atomic)']:>5} "
                       f"{counts['... of which .SYS scope']:>4} {counts['ATOMS (shared atomic)']:>5} {counts['VOTE']:>4}")
    with open(os.path.join(ROOT, "profiles", f"{prefix}_sass_summary.txt"), "w") as f:
        f.write("\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    main()
