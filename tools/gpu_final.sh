#!/bin/bash
# round-end style visit: parity tests, the bench line (with the CPU baseline), the other full-model configs
TAG=${1:-r01m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; head -c 330 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_bench.err
for c in 1 3; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config$c.json 2> gpurun_out/${TAG}_bench_config$c.err
  echo "config $c exit $?"; head -c 330 gpurun_out/${TAG}_bench_config$c.json; echo; tail -3 gpurun_out/${TAG}_bench_config$c.err
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/${TAG}_smoke.log
exit 0
