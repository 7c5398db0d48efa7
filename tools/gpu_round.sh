#!/bin/bash
# One gpurun call of the development loop on ONE GPU: smoke, sequence stress, bench (+ timeline), full GPU suite.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-dev}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== seq stress"; timeout 300 python tools/seq_stress.py 2>&1 | tail -4
echo "== full gpu suite" ; timeout 1800 python -m pytest tests -m gpu -q -rA 2>&1 | grep -v WARNING > gpurun_out/${TAG}_pytest_full.log; tail -3 gpurun_out/${TAG}_pytest_full.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 --timeline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("us/query", round(d["us_per_query"],2), "value M/s", round(d["value"]/1e6,1), "frac", round(d["roofline"]["frac"],3), "iso", round(d["roofline"]["kernel_ms_isolated"]*1e3,2),
      "| e2e bulk us", round(d["e2e"]["us_per_query"],1), "M/s", round(d["e2e"]["value"]/1e6,1), "single", round(d["e2e"]["single_call"]["us_per_query"],1),
      "python", round(d["e2e"]["python_search"]["us_per_query"],1), "kmers", round(d["e2e"]["kmers_path"]["us_per_query"],1),
      "| batch us/q", round(d["batched_64_queries_one_launch"]["us_per_query"],2), "| launches", d["gpu_launches"])
print("cpu", {k: (round(v["value"]) if isinstance(v, dict) and "value" in v else v) for k, v in d["cpu_baseline"].items() if k in ("value", "cores", "reference_python")})
PY
grep -v "^$" gpurun_out/${TAG}_bench.err | tail -28
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-400
