"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, total and average time per kernel,
and -- for the device-resident joins only (launches before the first k_probe_compact that covers a chunk of S, i.e. the
host-buffer leg, are found by grid-independent order: the first `joins` joins of `per_join` launches after generation) --
each kernel's share of the join's kernel time.  usage: launch_list_summary.py launches.csv joins"""
import collections, csv, io, sys
path, joins = sys.argv[1], int(sys.argv[2])
text = "".join(l for l in open(path) if l.startswith('"'))
rows = list(csv.DictReader(io.StringIO(text)))
def short(n):
    n = n.split("(")[0]
    return n.replace("void ", "").replace("hwbrj::", "")
names = [short(r["Kernel Name"]) for r in rows]
ns = [float(r["Metric Value"]) for r in rows]
# device-resident joins: a join starts with k_build_hist<...> that follows k_export_row / k_generate
starts = [i for i, n in enumerate(names) if n.startswith("k_build_hist") and (i == 0 or names[i - 1] in ("k_export_row", "k_generate", "k_copy8"))]
resident = set()
for j, st in enumerate(starts[:joins]):
    end = starts[j + 1] if j + 1 < len(starts) else len(names)
    resident.update(range(st, end))
tot = collections.OrderedDict()
for i, (n, t) in enumerate(zip(names, ns)):
    a = tot.setdefault(n, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t
    if i in resident: a[2] += t
res_total = sum(a[2] for a in tot.values())
print(f"{len(rows)} launches; device-resident joins found: {min(joins, len(starts))} ({len(resident)} launches, {res_total / 1e6 / max(min(joins, len(starts)), 1):.3f} ms of kernel time per join under ncu)")
print(f"{'kernel':44s}{'n':>5s}{'total ms':>11s}{'avg us':>11s}{'share of the resident joins':>30s}")
for n, (c, t, r) in tot.items():
    print(f"{n:44s}{c:5d}{t / 1e6:11.3f}{t / c / 1e3:11.1f}{(r / res_total if res_total else 0):30.3f}")
