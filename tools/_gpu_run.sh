mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "cuda_build or fold or lenet or transcript" > gpurun_out/s4_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest2.log
tail -5 gpurun_out/s4_pytest2.log
for b in 16 20 22 24 26; do timeout 120 python tools/microbench.py fold $b 20; done 2>&1 | tee gpurun_out/s4_fold_sweep.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s4_bench2.json 2> gpurun_out/s4_bench2.err
cut -c1-200 gpurun_out/s4_bench2.json; tail -3 gpurun_out/s4_bench2.err
