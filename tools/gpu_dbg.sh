#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q > gpurun_out/${1}_pytest_model.log 2>&1; grep -n "parity beyond\|AssertionError: (\|passed\|failed" gpurun_out/${1}_pytest_model.log | cut -c1-300
timeout 300 python tools/dbg_parity.py cityscapes 512 1024 32 1 > gpurun_out/${1}_dbg_city.txt 2>&1; grep "ctrl_out\|y_out \|attn_box" gpurun_out/${1}_dbg_city.txt | cut -c1-420
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${1}_bench.json 2> gpurun_out/${1}_bench.err
head -c 300 gpurun_out/${1}_bench.json
exit 0
