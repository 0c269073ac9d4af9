#!/bin/bash
# tile-order / dilate-kernel A/B on the default bench
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 200 --no-extras > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("%-22s value %.0f frames/s  ms/step %.4f" % ("$name", d["value"], d["ms_per_step"]))
except Exception as e:
    print("$name", "failed", e)
PY
}
run new A=1
run old CBINFER_DILATE_TILES=0
run new_scramble CBINFER_TILE_SCRAMBLE=1
run old_scramble CBINFER_DILATE_TILES=0 CBINFER_TILE_SCRAMBLE=1
run new2 A=1
run old2 CBINFER_DILATE_TILES=0
