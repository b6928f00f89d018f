#!/bin/sh
# Run ON THE GPU BOX (through gpurun): one `ncu --set full` capture per headline kernel, summaries exported as CSV so that
# the (large) .ncu-rep files need not travel.  usage: tools/ncu_capture.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
cap() {  # name regex skip count cmd...
    name=$1; rx=$2; skip=$3; cnt=$4; shift 4
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o /tmp/prof_$name "$@" > /tmp/ncu_$name.log 2>&1
    ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $OUT/ncu_${name}_${TAG}_raw.csv 2>/dev/null
    ncu -i /tmp/prof_$name.ncu-rep --page source --csv > /tmp/src_$name.csv 2>/dev/null
    # per-instruction page is large: keep the 150 hottest SASS lines by samples
    python3 - "$name" "$TAG" <<'PY'
import csv, sys
name, tag = sys.argv[1], sys.argv[2]
try:
    rows = list(csv.reader(open(f"/tmp/src_{name}.csv")))
    hdr = next(i for i, r in enumerate(rows) if any("Sampling" in c or "Samples" in c for c in r))
    head = rows[hdr]
    col = next(i for i, c in enumerate(head) if "Samples" in c or "Sampling" in c)
    body = [r for r in rows[hdr + 1:] if len(r) > col and r[col].replace(",", "").isdigit()]
    body.sort(key=lambda r: -int(r[col].replace(",", "")))
    with open(f"gpurun_out/ncu_{name}_{tag}_hot_sass.csv", "w", newline="") as f:
        w = csv.writer(f); w.writerow(head); w.writerows(body[:150])
except Exception as e:
    print("source page:", e)
PY
    echo "$name:"; tail -1 /tmp/ncu_$name.log
}
cap fold_tma k_round_quad_tma 1 1 python tools/microbench.py fold 24 2
cap fold_thin k_round_quad_thin 1 1 python tools/microbench.py fold 12 2
cap msm_small k_msm_small 1 1 python tools/microbench.py msm 12 12 2 1
cap msm_window k_msm_window 1 1 python tools/microbench.py msm 1 12 0 1
