#!/bin/bash
TAG=${1:-r01n}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err
echo "bench 2gpu exit $?"; head -c 400 gpurun_out/${TAG}_bench_2gpu.json; echo; tail -3 gpurun_out/${TAG}_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  tools/dp_step_check.py > gpurun_out/${TAG}_dp_step_check.txt 2>&1
echo "dp check exit $?"; tail -3 gpurun_out/${TAG}_dp_step_check.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref_2gpu.json 2> gpurun_out/${TAG}_bench_ref_2gpu.err
echo "ref arm exit $?"; head -c 300 gpurun_out/${TAG}_bench_ref_2gpu.json
exit 0
