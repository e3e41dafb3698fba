for D in ${@:-0 16 15 31}; do PVD_MLP_DIAG=$D timeout 120 python bench.py --workload mlp-hash --steps 20 --no-cpu-baseline --no-graph 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('diag $D teacher_fwd us', round(1000*d['kernel_ms']['teacher_fwd'],1))
    except Exception: pass
"; done
