#!/bin/bash
# One GPU-box pass: staged parity tests (separate processes so a faulting kernel cannot take the
# other stages down), the tcgen05 self-test, smoke, a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 300"
echo "== A: detect/compact/pool/fg/staged" ; timeout 900 $PYT tests/test_gpu_ops.py -m gpu -k "not conv_update" > gpurun_out/A.log 2>&1; echo "exit $?"; tail -3 gpurun_out/A.log
echo "== B: conv simt" ; timeout 900 $PYT tests/test_gpu_ops.py -m gpu -k "conv_update and simt" > gpurun_out/B.log 2>&1; echo "exit $?"; tail -3 gpurun_out/B.log
echo "== S: umma selftest" ; timeout 300 python tools/umma_selftest.py > gpurun_out/S.log 2>&1; echo "exit $?"; tail -25 gpurun_out/S.log
echo "== C: conv tc" ; timeout 900 $PYT tests/test_gpu_ops.py -m gpu -k "conv_update and not simt" > gpurun_out/C.log 2>&1; echo "exit $?"; tail -3 gpurun_out/C.log
echo "== D: modules" ; timeout 1200 $PYT tests/test_gpu_modules.py -m gpu > gpurun_out/D.log 2>&1; echo "exit $?"; tail -5 gpurun_out/D.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -3 gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
