#!/bin/bash
# round-2 multi-GPU pass: N = number of visible GPUs
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
if [ "$N" = "2" ]; then
  echo "== two-GPU split test"; timeout 600 python -m pytest -q --tb=short -p no:cacheprovider -x tests/test_gpu_split2.py -m gpu > gpurun_out/S2.log 2>&1; echo "exit $?"; tail -3 gpurun_out/S2.log
fi
echo "== suite $N gpus"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  benchmarks/multi_gpu_suite.py ${SUITE_ARGS:-} > gpurun_out/r02_suite_${N}gpu.jsonl 2> gpurun_out/suite${N}.err; echo "exit $?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/suite${N}.err | tail -5; cat gpurun_out/r02_suite_${N}gpu.jsonl
if [ -n "${E2E_BIND:-}" ]; then
  echo "== e2e with core binding"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    benchmarks/multi_gpu_suite.py --sections e2e --bind > gpurun_out/r02_suite_${N}gpu_bind.jsonl 2> gpurun_out/suite${N}b.err; echo "exit $?"
  cat gpurun_out/r02_suite_${N}gpu_bind.jsonl
fi
