#!/usr/bin/env python3
"""Summarise an ncu metrics pass over ONE proof (csv, --page raw style long format: one row per (launch, metric)) per kernel and per
kernel class of bench.py: launches, device time, DRAM bytes (read + write), mean SM throughput, registers.
usage: ncu_class_traffic.py pass.csv [out.json] [tag]
The class of a launch follows include/zkcnn_b200.h (ZK_PROF_*); sumcheck round kernels count as `fold` when they move at least
24 MiB of DRAM traffic and as `fold_small` otherwise (the library draws the same line at 32 MiB of algorithmic bytes)."""
import collections
import csv
import json
import sys

CLASS = {
    "k_round_quad_tma": "fold", "k_round_quad": "fold", "k_round_quad_thin": "fold", "k_round_cubic_tma": "fold", "k_round_cubic": "fold", "k_round_tail": "fold",
    "k_gate_items_p1": "gates", "k_gate_items_p2": "gates", "k_sum_partials": "gates",
    "k_msm_window": "msm", "k_msm_small": "msm", "k_msm_finish_rows": "msm", "k_msm_table_build": "msm", "k_msm_multiples_build": "msm", "k_msm_rowinfo": "msm",
    "k_half_tables": "tables", "k_beta_expand": "tables", "k_beta_outer": "tables", "k_phi_table": "tables",
    "k_dense_colsum": "dense", "k_colsum_finish": "dense", "k_dotprod_axpy": "dense", "k_dense_rowdot": "dense", "k_gather": "dense", "k_liu_scatter": "dense",
}


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    idi, ki, mi, vi = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    launches = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        d = launches.setdefault(r[idi], {"kernel": name})
        try:
            d[r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    per_kernel = collections.defaultdict(lambda: collections.defaultdict(float))
    per_class = collections.defaultdict(lambda: collections.defaultdict(float))
    for d in launches.values():
        b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        t = d.get("gpu__time_duration.sum", 0.0)
        cls = CLASS.get(d["kernel"], "other")
        if cls == "fold" and b < (24 << 20):
            cls = "fold_small"
        for tgt in (per_kernel[d["kernel"]], per_class[cls]):
            tgt["launches"] += 1
            tgt["dram_bytes"] += b
            tgt["ns"] += t
            tgt["sm_pct_x_ns"] += d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0) * t
        per_kernel[d["kernel"]]["regs"] = max(per_kernel[d["kernel"]]["regs"], d.get("launch__registers_per_thread", 0.0))
        per_kernel[d["kernel"]]["max_ns"] = max(per_kernel[d["kernel"]]["max_ns"], t)
    tot = sum(v["ns"] for v in per_kernel.values())
    print(f"{'kernel':28s} {'n':>6s} {'ms':>9s} {'share':>6s} {'max us':>9s} {'DRAM MB':>10s} {'GB/s':>8s} {'SM %':>6s} {'regs':>5s}")
    for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"{k:28s} {int(v['launches']):6d} {v['ns'] / 1e6:9.3f} {v['ns'] / tot:6.1%} {v['max_ns'] / 1e3:9.1f} {v['dram_bytes'] / 1e6:10.1f} "
              f"{v['dram_bytes'] / max(1.0, v['ns']):8.1f} {v['sm_pct_x_ns'] / max(1.0, v['ns']):6.1f} {int(v['regs']):5d}")
    print(f"total {tot / 1e6:.3f} ms over {len(launches)} launches (serialised, cold caches: compare shares)")
    print()
    out = {}
    for c, v in sorted(per_class.items(), key=lambda kv: -kv[1]["ns"]):
        print(f"class {c:12s} launches {int(v['launches']):6d}  {v['ns'] / 1e6:9.3f} ms  DRAM {v['dram_bytes'] / 1e9:8.3f} GB  {v['dram_bytes'] / max(1.0, v['ns']):8.1f} GB/s  SM {v['sm_pct_x_ns'] / max(1.0, v['ns']):5.1f} %")
        out["class_" + c] = int(v["dram_bytes"])
    if len(sys.argv) > 2:
        tag = sys.argv[3] if len(sys.argv) > 3 else ""
        try:
            old = json.load(open(sys.argv[2]))
        except Exception:
            old = {}
        if tag:
            old[tag] = out
        else:
            old.update(out)
        old.setdefault("source", {})[tag or "classes"] = f"{sys.argv[1]}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... over the launches of ONE proof; per-proof DRAM bytes per kernel class"
        json.dump(old, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
