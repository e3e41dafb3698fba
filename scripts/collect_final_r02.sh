#!/bin/bash
# Round-2 final single-GPU evidence: test suite, both bench arms (all workloads), ncu captures.  Outputs under gpurun_out/.
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/r02b_pytest_gpu.log 2>&1; tail -4 $O/r02b_pytest_gpu.log
timeout 900 python bench.py --impl reference > $O/r02b_bench_reference.json 2> $O/bench_ref_err.log; tail -2 $O/bench_ref_err.log
timeout 900 python bench.py > $O/r02b_bench_ours.json 2> $O/bench_ours_err.log; tail -2 $O/bench_ours_err.log
python - <<'PY'
import json
for f in ("r02b_bench_ours.json", "r02b_bench_reference.json"):
    try:
        j = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, j["ms_per_step"], {k: (v.get("ms_per_step"), (v.get("iteration") or {}).get("ms"), v.get("error")) for k, v in j["workloads"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/collect_r02.sh r02b > $O/collect.log 2>&1; tail -12 $O/collect.log
python tests/bench_inference.py > $O/r02b_inference_frame.json 2>$O/inf_err.log; tail -2 $O/inf_err.log
