#!/bin/bash
# round-2 pass B (1 GPU): planar-detect retest, rate-0 sweep rows, single-GPU rows of the multi-GPU suite
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 600 -x"
echo "== A: ops" ; timeout 1200 $PYT tests/test_gpu_ops.py -m gpu > gpurun_out/A.log 2>&1; echo "exit $?"; tail -3 gpurun_out/A.log
echo "== D: fg module" ; timeout 600 $PYT tests/test_gpu_modules.py -m gpu -k "fine_grained or split" > gpurun_out/D.log 2>&1; echo "exit $?"; tail -3 gpurun_out/D.log
echo "== sweep rate 0" ; timeout 600 python benchmarks/sweep_layers.py --rates 0,0.05 --layers scene_L2,pose_conv1_2,pose_conv4_2 > gpurun_out/r02_sweep_rate0.jsonl 2> gpurun_out/sweep.err; echo "exit $?"; tail -3 gpurun_out/sweep.err
echo "== suite 1 gpu" ; timeout 900 python benchmarks/multi_gpu_suite.py --split-rates 0.05,0.2,0.5,1.0 > gpurun_out/r02_suite_1gpu.jsonl 2> gpurun_out/suite1.err; echo "exit $?"; tail -5 gpurun_out/suite1.err; cat gpurun_out/r02_suite_1gpu.jsonl
