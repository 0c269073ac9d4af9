#!/bin/bash
# full GPU test-suite (staged) + compat shim tests + sweep rows that the tile policy / planar rule touch
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 600 -x"
echo "== C: compat shim" ; timeout 600 $PYT tests/test_gpu_compat_shim.py -m gpu > gpurun_out/C.log 2>&1; echo "exit $?"; tail -8 gpurun_out/C.log
echo "== A: ops" ; timeout 1200 $PYT tests/test_gpu_ops.py -m gpu > gpurun_out/A.log 2>&1; echo "exit $?"; tail -3 gpurun_out/A.log
echo "== D: modules" ; timeout 1500 $PYT tests/test_gpu_modules.py -m gpu > gpurun_out/D.log 2>&1; echo "exit $?"; tail -5 gpurun_out/D.log
echo "== T: tiles" ; timeout 900 $PYT tests/test_gpu_tiles.py -m gpu > gpurun_out/T.log 2>&1; echo "exit $?"; tail -3 gpurun_out/T.log
echo "== R: other gpu tests" ; timeout 1500 $PYT tests -m gpu --ignore tests/test_gpu_tiles.py --ignore tests/test_gpu_ops.py --ignore tests/test_gpu_modules.py --ignore tests/test_gpu_compat_shim.py > gpurun_out/R.log 2>&1; echo "exit $?"; tail -3 gpurun_out/R.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -2 gpurun_out/smoke.log
echo "== sweep" ; timeout 900 python benchmarks/sweep_layers.py --rates 0,0.01,0.05,0.2,1.0 > gpurun_out/r02_sweep_layers.jsonl 2> gpurun_out/sweep.err; echo "exit $?"; tail -3 gpurun_out/sweep.err
