#!/bin/bash
# One gpurun call of the development loop: smoke, targeted parity tests, bench (+ timeline), full GPU suite.
# usage: tools/gpu_round.sh <tag> [pytest -k expression for the targeted pass]
TAG=${1:-dev}
KEXPR=${2:-"exchange or solo_path or zero_copy or device_pointer or search_sequence or golden or config1"}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== seq stress"; timeout 300 python tools/seq_stress.py 2>&1 | tail -16
echo "== targeted tests" ; timeout 900 python -m pytest tests -m gpu -q -k "$KEXPR" > gpurun_out/${TAG}_pytest_targeted.log 2>&1; tail -12 gpurun_out/${TAG}_pytest_targeted.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 --timeline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("us/query", round(d["us_per_query"],2), "frac", round(d["roofline"]["frac"],3), "iso", round(d["roofline"]["kernel_ms_isolated"]*1e3,2),
      "| e2e bulk us", round(d["e2e"]["us_per_query"],1), "single", round(d["e2e"]["single_call"]["us_per_query"],1),
      "python", round(d["e2e"]["python_search"]["us_per_query"],1), "kmers", round(d["e2e"]["kmers_path"]["us_per_query"],1),
      "| batch us/q", round(d["batched_64_queries_one_launch"]["us_per_query"],2))
PY
tail -32 gpurun_out/${TAG}_bench.err
echo "== full gpu suite" ; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_full.log 2>&1; tail -15 gpurun_out/${TAG}_pytest_full.log
