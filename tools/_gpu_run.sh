mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s4_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest8.log
tail -4 gpurun_out/s4_pytest8.log
ZKH_TRACE=1 python tools/_trace_probe.py resident 6 2>&1 | grep -E "ZKH_TRACE|resident" | tail -3
