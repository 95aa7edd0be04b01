#!/bin/bash
TAG=${1:-r01g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest.log 2>&1
rc=$?
echo "pytest exit $rc" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
if [ $rc -ne 0 ]; then
  RA_CONV_PDL=0 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_pytest_nopdl.log 2>&1
  echo "pytest (no PDL) exit $?" >> gpurun_out/${TAG}_pytest_nopdl.log
  tail -6 gpurun_out/${TAG}_pytest_nopdl.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; head -c 330 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_bench.err
RA_CONV_PDL=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_nopdl.json 2> gpurun_out/${TAG}_bench_nopdl.err
echo "bench nopdl exit $?"; head -c 330 gpurun_out/${TAG}_bench_nopdl.json; echo
exit 0
