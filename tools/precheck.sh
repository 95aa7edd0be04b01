#!/bin/bash
# CPU gate before any gpurun call: build, host-flow dry run (every C-ABI call stubbed), quick CPU tests.
set -euo pipefail
cd "$(dirname "$0")/.."
make -s -j8 -C rec-attend-public_b200/csrc
python tools/dryrun_host.py > /tmp/dryrun.log 2>&1 || { tail -5 /tmp/dryrun.log; exit 1; }
python -m pytest tests/test_capi_load.py tests/test_params.py -q -x 2>&1 | tail -1
echo PRECHECK_OK
