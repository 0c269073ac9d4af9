#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 600 -x"
echo "== D: modules" ; timeout 1500 $PYT tests/test_gpu_modules.py -m gpu > gpurun_out/D.log 2>&1; echo "exit $?"; tail -8 gpurun_out/D.log
echo "== A: ops" ; timeout 1200 $PYT tests/test_gpu_ops.py -m gpu > gpurun_out/A.log 2>&1; echo "exit $?"; tail -3 gpurun_out/A.log
for v in 0 1; do
  echo "== CPM batch 1, CBINFER_FUSE_SMALL=$v"; CBINFER_FUSE_SMALL=$v timeout 600 python benchmarks/pose_cpm.py > gpurun_out/cpm_small$v.jsonl 2> gpurun_out/cpm.err; echo "exit $?"; cat gpurun_out/cpm_small$v.jsonl | cut -c1-700; tail -2 gpurun_out/cpm.err
done
echo "== CPM batch 8"; timeout 600 python benchmarks/pose_cpm.py --batch 8 > gpurun_out/cpm_b8.jsonl 2> gpurun_out/cpm.err; echo "exit $?"; cat gpurun_out/cpm_b8.jsonl | cut -c1-700
echo "== CPM batch 1, tile policy off (CBINFER_TILE_CLK only)"; CBINFER_TILES=0 timeout 600 python benchmarks/pose_cpm.py > gpurun_out/cpm_notiles.jsonl 2> gpurun_out/cpm.err; echo "exit $?"; cat gpurun_out/cpm_notiles.jsonl | cut -c1-700
