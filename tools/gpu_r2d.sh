#!/bin/bash
# round-2 pass D (1 GPU): tests with the lean dilate_tiles kernel, default bench, PDL variants, ncu of the new kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
PYT="python -m pytest -q --tb=short -p no:cacheprovider --timeout 600 -x"
echo "== T: tiles" ; timeout 900 $PYT tests/test_gpu_tiles.py -m gpu > gpurun_out/T.log 2>&1; echo "exit $?"; tail -5 gpurun_out/T.log
echo "== D: modules" ; timeout 1500 $PYT tests/test_gpu_modules.py tests/test_gpu_parity_baseline.py -m gpu > gpurun_out/D.log 2>&1; echo "exit $?"; tail -5 gpurun_out/D.log
echo "== bench (default, full)"; timeout 900 python bench.py > gpurun_out/r02_bench_full.json 2> gpurun_out/bench_full.err; echo "exit $?"; tail -3 gpurun_out/bench_full.err
for v in old pdl pdllate; do
  case $v in
    old) export CBINFER_DILATE_TILES=0; unset CBINFER_PDL CBINFER_LIB;;
    pdl) unset CBINFER_DILATE_TILES CBINFER_LIB; export CBINFER_PDL=1;;
    pdllate) unset CBINFER_DILATE_TILES; export CBINFER_PDL=1 CBINFER_LIB=$PWD/build/libcbinfer_pdllate.so;;
  esac
  echo "== bench variant $v"; timeout 600 python bench.py --steps 200 --no-extras > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "exit $?"
done
unset CBINFER_DILATE_TILES CBINFER_PDL CBINFER_LIB
python - <<PY
import json
for v in ("../gpurun_out/r02_bench_full", "bench_old", "bench_pdl", "bench_pdllate"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % v.replace("../gpurun_out/", "")).read().strip().splitlines()[-1])
        print("%-14s value %.0f frames/s  ms/step %.4f  e2e %.0f  u8 %.0f  labels %.0f  ceiling %.1f" % (v.split("/")[-1], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("e2e_u8_ingest", {}).get("value", 0), d.get("e2e_u8_ingest", {}).get("labels_out", {}).get("value", 0), d["e2e"].get("host_h2d_copy_only_gbs_per_gpu", 0)))
    except Exception as e:
        print(v, "parse failed", e)
PY
echo "== PDL tests"; CBINFER_PDL=1 timeout 900 $PYT tests/test_gpu_tiles.py tests/test_gpu_modules.py -m gpu > gpurun_out/P.log 2>&1; echo "exit $?"; tail -3 gpurun_out/P.log
echo "== ncu: the small HBM-side kernels at bench sizes"
CMD="python tools/aux_kernels.py"
$CMD > gpurun_out/aux_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"detect_planar|fg_detect|dilate_|detect_vec|detect_sparse|maxpool" -f -o gpurun_out/r02_aux $CMD > gpurun_out/ncu_aux.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/aux_plain.log
ncu -i gpurun_out/r02_aux.ncu-rep --page raw --csv > gpurun_out/r02_aux_raw.csv 2>/dev/null
ls -la gpurun_out/r02_aux.ncu-rep
