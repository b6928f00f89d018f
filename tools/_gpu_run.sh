mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cuda_build or fold or lenet or transcript" > gpurun_out/s4_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest2.log
tail -5 gpurun_out/s4_pytest2.log
TAG=s4c
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_round_quad_tma -s 1 -c 1 -f -o /tmp/prof_fold python tools/microbench.py fold 24 2 > /tmp/ncu_fold.log 2>&1
ncu -i /tmp/prof_fold.ncu-rep --page raw --csv > gpurun_out/ncu_fold_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_fold.ncu-rep --page source --csv > gpurun_out/ncu_fold_${TAG}_source.csv 2>/dev/null
tail -2 /tmp/ncu_fold.log
