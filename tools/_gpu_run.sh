mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "cuda_build or fold or lenet or transcript" > gpurun_out/s4_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest2.log
tail -3 gpurun_out/s4_pytest2.log
for thin in 16384 0; do
for b in 4 10 14 16; do
ZK_THIN_MAX_PAIRS=$thin timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_round_quad python tools/microbench.py fold $b 5 2>/dev/null | grep k_round | awk -F'","' -v t=$thin -v b=$b '{print "thin="t" bits="b" "$5" "$NF}' | tail -1
done; done 2>&1 | tee gpurun_out/s4_small_ncu.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s4_bench2.json 2> gpurun_out/s4_bench2.err
cut -c1-200 gpurun_out/s4_bench2.json; tail -3 gpurun_out/s4_bench2.err
