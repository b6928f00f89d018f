#!/bin/sh
# Run ON THE GPU BOX (through gpurun): ncu captures of round 2.  `--set full` for the two streaming round kernels (micro-benchmarks, one
# launch each); the speed-of-light / memory / occupancy / launch sections for one in-proof launch of every other kernel of SURVEY section 8
# (inside the NVTX range "proof" of the probe scripts).  Summaries are exported as CSV so that the .ncu-rep files need not travel.
OUT=gpurun_out
mkdir -p $OUT
full() {  # name regex cmd...
    name=$1; rx=$2; shift 2
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -f -o /tmp/prof_$name "$@" > /tmp/ncu_$name.log 2>&1
    ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $OUT/r02_ncu_${name}_raw.csv 2>/dev/null
    ncu -i /tmp/prof_$name.ncu-rep --page source --csv > /tmp/src_$name.csv 2>/dev/null
    python3 - "$name" <<'PY'
import csv, sys
name = sys.argv[1]
try:
    rows = list(csv.reader(open(f"/tmp/src_{name}.csv")))
    hdr = next(i for i, r in enumerate(rows) if any("Sampling" in c or "Samples" in c for c in r))
    head = rows[hdr]
    col = next(i for i, c in enumerate(head) if "Samples" in c or "Sampling" in c)
    body = [r for r in rows[hdr + 1:] if len(r) > col and r[col].replace(",", "").isdigit()]
    body.sort(key=lambda r: -int(r[col].replace(",", "")))
    with open(f"gpurun_out/r02_ncu_{name}_hot_sass.csv", "w", newline="") as f:
        w = csv.writer(f); w.writerow(head); w.writerows(body[:120])
except Exception as e:
    print("source page:", e)
PY
    echo "$name: $(tail -1 /tmp/ncu_$name.log)"
}
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats --section ComputeWorkloadAnalysis"
inproof() {  # name regex skip cmd...
    name=$1; rx=$2; skip=$3; shift 3
    PROBE_NVTX=1 timeout 600 ncu $SECT --clock-control none --nvtx --nvtx-include "proof/" -k regex:$rx -s $skip -c 1 -f -o /tmp/prof_$name "$@" > /tmp/ncu_$name.log 2>&1
    ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $OUT/r02_ncu_${name}_raw.csv 2>/dev/null
    echo "$name: $(tail -1 /tmp/ncu_$name.log)"
}
full fold_tma k_round_quad_tma python tools/microbench.py fold 24 3
full cubic_tma k_round_cubic_tma python tools/microbench.py cubic 24 3
# FFT-convolution path (vgg11, two pictures): K4b, K5b, K5 dense passes
inproof dense_colsum 'k_dense_colsum' 1 python tools/probe_fft_path.py vgg11 2 1
inproof dotprod_axpy 'k_dotprod_axpy' 1 python tools/probe_fft_path.py vgg11 2 1
inproof dense_rowdot 'k_dense_rowdot' 1 python tools/probe_fft_path.py vgg11 2 1
# vgg11, one picture: K3, K4, K6, fused tail, K9, f-1
inproof half_tables 'k_half_tables' 60 python tools/probe_proofs.py resident 2
inproof gate_items_p1 'k_gate_items_p1' 30 python tools/probe_proofs.py resident 2
inproof liu_scatter 'k_liu_scatter' 20 python tools/probe_proofs.py resident 2
inproof round_tail 'k_round_tail' 90 python tools/probe_proofs.py resident 2
inproof msm_window_opening 'k_msm_window' 3 python tools/probe_proofs.py resident 2
inproof msm_finish_rows 'k_msm_finish_rows' 3 python tools/probe_proofs.py resident 2
python3 tools/ncu_summary_r02.py $OUT > $OUT/r02_ncu_summary.txt 2>&1; cat $OUT/r02_ncu_summary.txt
