#!/bin/bash
# bench value under a list of environment settings, same box, same minute.  usage: tools/gpu_knobs.sh "A=1 B=0" "C=1" ...
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
i=0
for cfg in "" "$@"; do
  i=$((i+1))
  env $cfg timeout 600 python bench.py --steps 400 --warmup 10 --no-extras > gpurun_out/knob_$i.json 2> gpurun_out/knob_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/knob_$i.json").read().strip().splitlines()[-1])
    print("[%s]: value %.0f frames/s  ms/step %.4f  u8 %.0f" % ("$cfg", d["value"], d["ms_per_step"], d.get("e2e_u8_ingest", {}).get("value", 0)))
except Exception as e:
    print("bench parse failed [$cfg]", e); print(open("gpurun_out/knob_$i.err").read()[-1500:])
PY
done
