"""Per-source-line executed-instruction breakdown of one kernel in an .ncu-rep (needs -lineinfo + --import-source)."""
import collections, csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "-k", f"regex:{pat}", "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
agg, total, hdr, cur, fn = collections.Counter(), 0, None, None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Function Name":
        fn = r[1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr):
        try:
            line = int(r[0]); ie = int(r[hdr.index("Instructions Executed")])
        except Exception:
            continue
        agg[(cur, line, r[1].strip()[:100])] += ie; total += ie
print(fn); print("total warp instructions", total)
for (f, l, s), v in agg.most_common(top):
    print(f"{v / total * 100:5.1f}%  {f}:{l}  {s}")
