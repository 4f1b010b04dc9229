"""Regenerates tests/golden/golden_results.json from the reference's own published measurement data
(/root/reference/measurements/data/pkl/*.pkl, written by measurements/run.py:414-425) and
tests/golden/hash_kat.json / bitmap_kat.json from the compiled reference (oracle/_ref/libref.so).

Runs only in the authoring container (needs /root/reference and pandas); the JSON files are committed and are
what the tests read.  Usage:  python tests/golden/make_golden.py
"""
import glob
import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
PKL = "/root/reference/measurements/data/pkl"


def mine_pkl():
    rows = {}
    nrows = 0
    for f in sorted(glob.glob(os.path.join(PKL, "*.pkl"))):
        if "hash_functions" in f:
            continue
        df = pd.read_pickle(f)
        need = {"r-size", "s-size", "s-sel", "bloom-filter", "bloom-hashes", "bloom-size", "bloom-block-size",
                "filtered", "out-tuples"}
        if not need.issubset(df.columns):
            continue
        for rec in df[list(need)].itertuples(index=False):
            d = dict(zip(list(need), rec))
            if pd.isna(d["out-tuples"]) or (d["bloom-filter"] != "no" and pd.isna(d["filtered"])):
                continue  # failed/timeout runs in the published data
            nrows += 1
            bf = d["bloom-filter"]
            if bf == "no":
                key = (int(d["r-size"]), int(d["s-size"]), float(d["s-sel"]), "no", 0, 0, 0)
                val = (-1, int(d["out-tuples"]))
            else:
                key = (int(d["r-size"]), int(d["s-size"]), float(d["s-sel"]), bf, int(d["bloom-size"]),
                       int(d["bloom-hashes"]), int(d["bloom-block-size"]))
                val = (int(d["filtered"]), int(d["out-tuples"]))
            if key in rows and rows[key][:2] != val:
                raise SystemExit(f"inconsistent golden for {key}: {rows[key]} vs {val} in {f}")
            rows[key] = val + (rows.get(key, (0, 0, 0))[2] + 1,)
    out = [{"r": k[0], "s": k[1], "q": k[2], "bloom": k[3], "m": k[4], "k": k[5], "B": k[6], "filtered": v[0],
            "matches": v[1], "runs": v[2]} for k, v in sorted(rows.items())]
    return out, nrows


def hash_kat():
    import oracle
    keys = [0, 1, 2, 1000, 128000000, 128000001, 2147483647, -1, -2147483648, 77, -77, 0x80, 0x8000, 0x800000,
            0x7F7F7F7F, -0x7F7F7F7F, 305419896]
    seeds = [42, 0, 1, 0xDEADBEEF, 817263]
    rows = []
    for seed in seeds:
        for key in keys:
            rows.append({"seed": seed, "key": key, "hashes": [oracle.ref_hash(w, seed, key) for w in range(10)]})
    return {"order": oracle.HASH_NAMES, "rows": rows}


def bitmap_kat():
    """m=1024, seed 42, insert keys 1..64, count how many of keys 65..1064 pass."""
    import oracle
    R = np.zeros(64, dtype=oracle.TUPLE)
    R["key"] = np.arange(1, 65)
    S = np.zeros(1000, dtype=oracle.TUPLE)
    S["key"] = np.arange(65, 1065)
    rows = []
    for variant, k, B in [(0, 1, 512), (0, 3, 512), (1, 3, 64), (1, 1, 512), (1, 8, 256), (1, 2, 8), (0, 8, 512),
                          (1, 5, 1024)]:
        bm = oracle.ref_bloom_build(R, variant, 1024, k, B)
        n = oracle.ref_bloom_count(bm, S, variant, 1024, k, B)
        rows.append({"variant": variant, "k": k, "B": B, "m": 1024, "bitmap_hex": bm.tobytes().hex(), "pass": int(n)})
    return rows


if __name__ == "__main__":
    gold, n = mine_pkl()
    json.dump({"source": "measurements/data/pkl/*.pkl (reference repository)", "pkl_rows": n, "configs": gold},
              open(os.path.join(HERE, "golden_results.json"), "w"), indent=0)
    print(f"{len(gold)} distinct configs from {n} published rows")
    json.dump(hash_kat(), open(os.path.join(HERE, "hash_kat.json"), "w"))
    json.dump(bitmap_kat(), open(os.path.join(HERE, "bitmap_kat.json"), "w"), indent=1)
