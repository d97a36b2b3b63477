#!/bin/bash
# One GPU-box visit: engine self-test, parity tests, bf16-vs-fp32 error report, benches.  Logs into gpurun_out/.
TAG=${1:-r1}
mkdir -p gpurun_out
python tools/tc_selftest.py > gpurun_out/${TAG}_selftest.log 2>&1; echo "selftest rc=$?"; grep -c OK gpurun_out/${TAG}_selftest.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/debug_bf16.py mlp_full > gpurun_out/${TAG}_bf16err.log 2>&1; echo "bf16err rc=$?"; head -n 4 gpurun_out/${TAG}_bf16err.log
timeout 300 python tools/profile_step.py --rays 16384 --precision bf16 2>&1 | tail -n 1
timeout 600 python bench.py --precision bf16 --rays 65536 --steps 3 --warmup 3 > gpurun_out/${TAG}_bench65536_bf16.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench65536_bf16.log
