#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 600 -x"
echo "== A: ops (detect)" ; timeout 1200 $PYT tests/test_gpu_ops.py -m gpu -k "detect or kat" > gpurun_out/A.log 2>&1; echo "exit $?"; tail -3 gpurun_out/A.log
echo "== D: u8 + modules" ; timeout 1500 $PYT tests/test_gpu_modules.py -m gpu -k "uint8 or detect_input or scene" > gpurun_out/D.log 2>&1; echo "exit $?"; tail -3 gpurun_out/D.log
echo "== in-step timing"; timeout 600 python tools/instep_timing.py > gpurun_out/r02_instep_timing.txt 2>&1; echo "exit $?"; cat gpurun_out/r02_instep_timing.txt | tail -12
for i in 1 2; do timeout 600 python bench.py --steps 200 --no-extras > gpurun_out/ab_l$i.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/ab_l$i.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], 'u8', d['e2e_u8_ingest']['value'])"; done
