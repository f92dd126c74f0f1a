#!/usr/bin/env python
"""profiles/step_traffic.json from an ncu launch list of `bench.py --steps 1 ...` captured with
--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum: per kernel (bench.py's
names) the DRAM bytes and time of the LAST compute step (the launches after the last L2-flush memset).

usage: python scripts/step_traffic.py <launches.csv> <reads> [out.json]
"""
import csv
import json
import re
import sys


def bench_name(fn):
    m = re.search(r"(\w+)_kernel(<[^>]*>)?", fn)
    if not m:
        return None
    base, targs = m.group(1), m.group(2) or ""
    if base == "k1":
        return "k1_prefix" if re.search(r",\s*(1|true)\s*>$", targs) else "k1"
    if base.startswith("scan"):
        return "scan"
    return {"alpha_len": "alpha"}.get(base, base)


def main():
    path, reads = sys.argv[1], int(sys.argv[2])
    out = sys.argv[3] if len(sys.argv) > 3 else "profiles/step_traffic.json"
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    launches, order = {}, []
    for r in rows:
        i = int(r["ID"])
        if i not in launches:
            launches[i] = {"fn": r["Kernel Name"]}
            order.append(i)
        launches[i][r["Metric Name"]] = float(r["Metric Value"])
    flush = [i for i in order if "elementwise_kernel" in launches[i]["fn"]]
    start = flush[-1] if flush else -1
    kernels = {}
    for i in order:
        if i <= start:
            continue
        name = bench_name(launches[i]["fn"])
        if name is None:
            continue
        k = kernels.setdefault(name, {"launches": 0, "ms": 0.0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0})
        k["launches"] += 1
        k["ms"] += launches[i].get("gpu__time_duration.sum", 0.0) / 1e6
        k["dram_bytes_read"] += launches[i].get("dram__bytes_read.sum", 0.0)
        k["dram_bytes_write"] += launches[i].get("dram__bytes_write.sum", 0.0)
    total = sum(k["dram_bytes_read"] + k["dram_bytes_write"] for k in kernels.values())
    json.dump({"reads": reads, "source": path, "note": "per-launch values are cold-cache and serialised by ncu: shares, not absolutes",
               "total_dram_bytes": total, "total_ms": sum(k["ms"] for k in kernels.values()), "kernels": kernels},
              open(out, "w"), indent=1)
    print(json.dumps({"total_dram_GB": total / 1e9, "kernels": {k: round(v["ms"], 3) for k, v in kernels.items()}}))


if __name__ == "__main__":
    main()
