#!/bin/bash
# final pass without the ncu captures: the whole GPU test-suite as the driver runs it, smoke, both bench arms
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== pytest -m gpu"; ( time timeout 1800 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/final_pytest.log 2>&1; echo "exit $?"; tail -6 gpurun_out/final_pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -2 gpurun_out/smoke.log
echo "== bench (default)"; ( time python bench.py ) > gpurun_out/r02_bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; grep real gpurun_out/bench_full.err
echo "== bench --impl reference"; ( time python bench.py --impl reference ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/bench_ref.err; echo "exit $?"; grep real gpurun_out/bench_ref.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_full.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "parity_max_rel")}, d["e2e"]["value"], d["e2e_u8_ingest"]["value"], d["roofline"]["frac"], d["clocks"])
r = json.loads(open("gpurun_out/r02_bench_reference.json").read().strip().splitlines()[-1])
print({k: r.get(k) for k in ("impl", "value", "ms_per_step")})
PY
