#!/bin/bash
# tests + bench + conv fixed-cost diagnostics (timeline of the small layers)
TAG=${1:-r01c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; head -c 600 gpurun_out/${TAG}_bench.json; echo
RA_ONLY_TMA=1 timeout 300 python tools/conv_timeline.py attn_L0 attn_L4 dcnn_L0 dcnn_L1 dcnn_L6 ctrl_L7 ctrl_L1 tiny > gpurun_out/${TAG}_timeline.txt 2>&1
timeout 300 python tools/bench_conv_fixed.py > gpurun_out/${TAG}_conv_fixed.txt 2>&1
tail -12 gpurun_out/${TAG}_conv_fixed.txt
exit 0
