mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest3.log
tail -4 gpurun_out/s4_pytest3.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s4_bench3.json 2> gpurun_out/s4_bench3.err
cut -c1-200 gpurun_out/s4_bench3.json; tail -3 gpurun_out/s4_bench3.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --round-by-round > gpurun_out/s4_bench3_rbr.json 2> gpurun_out/s4_bench3_rbr.err
cut -c1-200 gpurun_out/s4_bench3_rbr.json
