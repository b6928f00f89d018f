#!/bin/sh
# Run ON THE GPU BOX (through gpurun): ncu captures of the MSM kernels as they are at the end of round 2 (one in-proof launch each, inside the
# NVTX range "proof" of tools/probe_proofs.py), and the launch list of one whole vgg11 proof.  Summaries are exported as CSV.
OUT=gpurun_out
mkdir -p $OUT
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats --section ComputeWorkloadAnalysis"
inproof() {  # name regex skip cmd...
    name=$1; rx=$2; skip=$3; shift 3
    PROBE_NVTX=1 timeout 600 ncu $SECT --clock-control none --nvtx --nvtx-include "proof/" -k regex:$rx -s $skip -c 1 -f -o /tmp/prof_$name "$@" > /tmp/ncu_$name.log 2>&1
    ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > $OUT/r02b_ncu_${name}_raw.csv 2>/dev/null
    echo "$name: $(tail -1 /tmp/ncu_$name.log)"
}
inproof msm_small 'k_msm_small' 0 python tools/probe_proofs.py resident 2
inproof msm_bucket_fill 'k_msm_bucket_fill' 0 python tools/probe_proofs.py resident 2
inproof msm_bucket_merge 'k_msm_bucket_merge' 0 python tools/probe_proofs.py resident 2
inproof msm_bucket_reduce 'k_msm_bucket_reduce' 0 python tools/probe_proofs.py resident 2
inproof msm_finish_rows 'k_msm_finish_rows' 0 python tools/probe_proofs.py resident 2
inproof msm_multiples_build 'k_msm_multiples_build' 0 python tools/probe_proofs.py resident 2
# launch list of ONE proof (the last of three)
PROBE_NVTX=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "proof/" --csv --log-file $OUT/r02b_launches_vgg11_one_proof.csv python tools/probe_proofs.py resident 3 > /tmp/launches.log 2>&1
python3 tools/launch_summary.py $OUT/r02b_launches_vgg11_one_proof.csv > $OUT/r02b_launches_summary.txt 2>&1; head -40 $OUT/r02b_launches_summary.txt
sed 's/r02_ncu_/r02b_ncu_/' tools/ncu_summary_r02.py > /tmp/ncu_summary_r02b.py; python3 /tmp/ncu_summary_r02b.py $OUT > $OUT/r02b_ncu_summary.txt 2>&1; cat $OUT/r02b_ncu_summary.txt
