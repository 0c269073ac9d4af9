#!/bin/bash
# final pass of a round (1 GPU): the whole GPU test-suite as the driver runs it, smoke, the default bench (full JSON
# line), ncu launch list + --set full of one steady-state step
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "exit $?"; tail -4 gpurun_out/final_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -2 gpurun_out/smoke.log
echo "== bench (default)"; ( time python bench.py ) > gpurun_out/r02_bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; grep real gpurun_out/bench_full.err
echo "== bench --impl reference"; ( time python bench.py --impl reference ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_ref.err; echo "exit $?"; grep real gpurun_out/bench_ref.err
bash tools/gpu_profile.sh
if [ -f build/libcbinfer_trace.so ]; then
  echo "== tile trace"; CBINFER_LIB=$PWD/build/libcbinfer_trace.so timeout 300 python tools/tile_trace.py > gpurun_out/r02_tile_trace.txt 2>&1; echo "exit $?"
fi
