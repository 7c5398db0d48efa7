#!/bin/bash
TAG=${1:-dev}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== exchange tests" ; timeout 600 python -m pytest tests -m gpu -x -q -k "exchange or solo_path or zero_copy" 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_targeted.log
echo "== bench default" ; timeout 600 python bench.py --steps 20 --warmup 5 --timeline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -30 gpurun_out/${TAG}_bench.err
for v in "n_stages=8" "n_stages=6" "pool_pct=12"; do
  echo "== bench $v"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt $v > gpurun_out/${TAG}_bench_${v}.json 2> gpurun_out/${TAG}_bench_${v}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${v}.json"))
print("$v", "us/query", round(d["us_per_query"],2), "frac", round(d["roofline"]["frac"],3), "iso", round(d["roofline"]["kernel_ms_isolated"]*1e3,2), "e2e us", round(d["e2e"]["us_per_query"],1))
PY
done
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("default", "us/query", round(d["us_per_query"],2), "frac", round(d["roofline"]["frac"],3), "iso", round(d["roofline"]["kernel_ms_isolated"]*1e3,2), "e2e us", round(d["e2e"]["us_per_query"],1), "batch us/q", round(d["batched_64_queries_one_launch"]["us_per_query"],2))
PY
