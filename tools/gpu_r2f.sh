#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 600 -x"
echo "== T: tiles" ; timeout 900 $PYT tests/test_gpu_tiles.py -m gpu > gpurun_out/T.log 2>&1; echo "exit $?"; tail -3 gpurun_out/T.log
echo "== tile policy"; timeout 600 python tools/tile_bench.py --set policy > gpurun_out/tile_policy.log 2>&1; echo "exit $?"; cat gpurun_out/tile_policy.log
echo "== planar nw sweep"
for nw in auto 1 2 4 8; do
  if [ $nw = auto ]; then unset CBINFER_PLANAR_NW; else export CBINFER_PLANAR_NW=$nw; fi
  timeout 300 python tools/planar_bench.py 2>&1 | tail -6
done
unset CBINFER_PLANAR_NW
echo "== CPM (config 4)"; timeout 600 python benchmarks/pose_cpm.py > gpurun_out/r02_pose_cpm.jsonl 2> gpurun_out/cpm.err; echo "exit $?"; cat gpurun_out/r02_pose_cpm.jsonl; tail -3 gpurun_out/cpm.err
