#!/bin/bash
# ncu passes (one GPU, each after the same command ran clean).
#   usage: gpu_profile.sh [list] [name:regex:skip:count ...]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 3 --warmup 3 --no-extras ${BENCH_ARGS:-}"
for spec in "$@"; do
  if [ "$spec" = "list" ]; then
    $CMD > gpurun_out/plain.log 2>&1 &&
    ncu --metrics gpu__time_duration.sum --clock-control none -s ${LIST_SKIP:-0} -c ${LIST_COUNT:-600} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
    echo "list exit $?"
  else
    IFS=: read name regex skip cnt <<< "$spec"
    $CMD > gpurun_out/plain_$name.log 2>&1 &&
    ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $cnt -f -o gpurun_out/prof_$name $CMD > gpurun_out/ncu_$name.log 2>&1
    echo "$name exit $?"
  fi
done
