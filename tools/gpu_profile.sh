#!/bin/bash
# ncu passes on one steady-state step of the bench (one GPU; each after the same command ran clean).
# Outputs: gpurun_out/r02_launches.csv (per-launch device times), gpurun_out/r02_full_raw.csv
# (--set full, raw page of every launch of one step), gpurun_out/r02_full.ncu-rep if small enough.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --profile 2 --steps 3 --warmup 5 ${BENCH_ARGS:-}"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list exit $?"
CMD1="python bench.py --profile 1 --steps 3 --warmup 5 ${BENCH_ARGS:-}"
$CMD1 > gpurun_out/plain1.log 2>&1 &&
ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r02_full $CMD1 > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2> gpurun_out/ncu_export.log
ncu -i gpurun_out/r02_full.ncu-rep --page details --csv > gpurun_out/r02_full_details.csv 2>> gpurun_out/ncu_export.log
ls -la gpurun_out/r02_full.ncu-rep
SZ=$(stat -c %s gpurun_out/r02_full.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 45000000 ]; then rm -f gpurun_out/r02_full.ncu-rep; echo "report too large, csv pages kept"; fi
tail -n 3 gpurun_out/ncu_list.log; tail -n 3 gpurun_out/ncu_full.log
