cd "$(dirname "$0")/.."
timeout 300 python -m pytest -q --tb=short -p no:cacheprovider -x tests/test_gpu_ops.py -m gpu -k "cta_pair" 2>&1 | tail -2
for pm in 0 64 16; do echo "== CBINFER_PAIR_MIN=$pm"; CBINFER_PAIR_MIN=$pm python benchmarks/sweep_layers.py --layers scene_L3 --rates 0.05,0.1,0.2,0.5,1.0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print(r['dtype'], r['rate'], {k:v for k,v in r.items() if k.startswith('cg')})"; done
