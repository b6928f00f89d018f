mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s4_bench4.json 2> gpurun_out/s4_bench4.err
cut -c1-300 gpurun_out/s4_bench4.json; tail -3 gpurun_out/s4_bench4.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-prefetch --round-by-round > gpurun_out/s4_bench4_rbr.json 2> gpurun_out/s4_bench4_rbr.err
cut -c1-300 gpurun_out/s4_bench4_rbr.json
