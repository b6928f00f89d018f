timeout 600 python -m pytest tests -m gpu -x -q -k "msm or hyrax or lenet_syn_p1_seed3_realgens" 2>&1 | tail -2
python tools/microbench.py msm 12 12 2 3
ZK_TRACE=1 python tools/_trace_probe.py resident 2 2> /tmp/t2.txt; ZK_TRACE=1 ZKH_TRACE=1 python tools/_trace_probe.py resident 12 2> /tmp/t12.txt
python tools/_trace_diff.py /tmp/t2.txt /tmp/t12.txt 10
grep ZKH_TRACE /tmp/t12.txt | tail -2; grep resident /tmp/t12.txt | tail -2
