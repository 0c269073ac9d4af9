#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 600 -x"
echo "== C: compat shim" ; timeout 600 $PYT tests/test_gpu_compat_shim.py -m gpu > gpurun_out/C.log 2>&1; echo "exit $?"; tail -4 gpurun_out/C.log
echo "== A: ops (detect)" ; timeout 1200 $PYT tests/test_gpu_ops.py -m gpu -k "detect" > gpurun_out/A.log 2>&1; echo "exit $?"; tail -3 gpurun_out/A.log
echo "== planar bench"; timeout 300 python tools/planar_bench.py 2>&1 | tail -6
echo "== in-step timing"; timeout 600 python tools/instep_timing.py > gpurun_out/r02_instep_timing.txt 2>&1; echo "exit $?"; cat gpurun_out/r02_instep_timing.txt | tail -20
echo "== in-step timing, lean dilation"; CBINFER_DILATE_TILES=1 timeout 600 python tools/instep_timing.py > gpurun_out/r02_instep_timing_lean.txt 2>&1; echo "exit $?"; cat gpurun_out/r02_instep_timing_lean.txt | tail -14
