#!/bin/bash
# One GPU-box pass: staged parity tests (separate processes so a faulting kernel cannot take the
# other stages down), smoke, short benches.  Logs land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 600 -x"
echo "== T: tiled contraction" ; timeout 900 $PYT tests/test_gpu_tiles.py -m gpu > gpurun_out/T.log 2>&1; echo "exit $?"; tail -15 gpurun_out/T.log
echo "== A: ops" ; timeout 1200 $PYT tests/test_gpu_ops.py -m gpu > gpurun_out/A.log 2>&1; echo "exit $?"; tail -5 gpurun_out/A.log
echo "== D: modules" ; timeout 1500 $PYT tests/test_gpu_modules.py -m gpu > gpurun_out/D.log 2>&1; echo "exit $?"; tail -8 gpurun_out/D.log
echo "== R: other gpu tests" ; timeout 1500 $PYT tests -m gpu --ignore tests/test_gpu_tiles.py --ignore tests/test_gpu_ops.py --ignore tests/test_gpu_modules.py > gpurun_out/R.log 2>&1; echo "exit $?"; tail -5 gpurun_out/R.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -3 gpurun_out/smoke.log
for t in 0 1; do
  echo "== bench tiles=$t" ; CBINFER_TILES=$t timeout 900 python bench.py --steps 200 --warmup 10 --no-extras > gpurun_out/bench_tiles$t.json 2> gpurun_out/bench_tiles$t.err; echo "exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_tiles$t.json").read().strip().splitlines()[-1])
    print("tiles=$t value %.0f frames/s  ms/step %.4f  e2e %.0f  u8 %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("e2e_u8_ingest", {}).get("value", 0)))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/bench_tiles$t.err").read()[-2000:])
PY
done
