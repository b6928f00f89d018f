timeout 600 python -m pytest tests -m gpu -x -q -k "msm or hyrax or lenet_syn_p1_seed3_realgens" 2>&1 | tail -2
python tools/microbench.py msm 12 12 2 3
ZKH_TRACE=1 python tools/_trace_probe.py resident 2>&1 | grep -E "ZKH_TRACE|resident" | tail -3
