#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 600 -x"
echo "== A: ops" ; timeout 1200 $PYT tests/test_gpu_ops.py -m gpu > gpurun_out/A.log 2>&1; echo "exit $?"; tail -3 gpurun_out/A.log
echo "== T: tiles" ; timeout 900 $PYT tests/test_gpu_tiles.py -m gpu > gpurun_out/T.log 2>&1; echo "exit $?"; tail -3 gpurun_out/T.log
echo "== D: modules" ; timeout 1500 $PYT tests/test_gpu_modules.py tests/test_gpu_parity_baseline.py -m gpu > gpurun_out/D.log 2>&1; echo "exit $?"; tail -3 gpurun_out/D.log
echo "== in-step timing"; timeout 600 python tools/instep_timing.py > gpurun_out/r02_instep_timing.txt 2>&1; echo "exit $?"; cat gpurun_out/r02_instep_timing.txt | tail -12
for i in 1 2; do timeout 600 python bench.py --steps 200 --no-extras > gpurun_out/ab_k$i.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/ab_k$i.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], 'u8', d['e2e_u8_ingest']['value'])"; done
