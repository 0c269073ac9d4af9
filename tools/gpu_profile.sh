#!/bin/bash
# ncu passes (one GPU, after the same command ran clean): launch list + full capture of the top kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 3 --warmup 3 --no-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${1:-conv_umma}" -s ${2:-12} -c ${3:-4} -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
tail -3 gpurun_out/ncu_full.log
