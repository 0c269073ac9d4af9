#!/bin/bash
# quick A/B of bench variants in one GPU call: each line = extra args
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/variants.txt
while IFS= read -r args; do
  [ -z "$args" ] && continue
  echo "## $args" >> gpurun_out/variants.txt
  python bench.py --steps 40 --warmup 5 --no-extras $args 2>> gpurun_out/variants.err | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('value %.0f fps  ms/step %.4f  e2e %.0f  counts %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('changed_pixels_last_frame')))
" >> gpurun_out/variants.txt
done
cat gpurun_out/variants.txt
