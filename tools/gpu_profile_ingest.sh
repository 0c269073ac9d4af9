#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python tools/ingest_probe.py > gpurun_out/ingest_plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none -f -o gpurun_out/r02_ingest python tools/ingest_probe.py > gpurun_out/ncu_ingest.log 2>&1
echo "ingest ncu exit $?"; tail -2 gpurun_out/ingest_plain.log
python tools/ncu_summary.py gpurun_out/r02_ingest.ncu-rep > gpurun_out/r02_ncu_ingest.txt 2>&1
