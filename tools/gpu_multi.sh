#!/bin/bash
# multi-GPU round: exchange parity tests (one process / one process per shard over CUDA IPC) + the driver's SCALE commands
# usage: tools/gpu_multi.sh <tag> "<list of N>"
TAG=${1:-multi}; NS=${2:-"2"}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== exchange / sharded tests"; timeout 900 python -m pytest tests -m gpu -q -rA -k "exchange or sharded or golden or config1" 2>&1 | grep -v WARNING > gpurun_out/${TAG}_pytest_multi.log; tail -4 gpurun_out/${TAG}_pytest_multi.log
for N in $NS; do
  echo "== bench N=$N"
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_bench_n$N.json") if l.startswith("{")][-1])
    print("N=$N value %.1f M/s  us/query %.2f  e2e %.1f M/s (%.1f us/query)  per-rank ms/step %s  wait_us %s" % (d["value"]/1e6, d["us_per_query"], d["e2e"]["value"]/1e6, d["e2e"]["us_per_query"], [round(x,3) for x in d["per_rank_ms_per_step"]], [round(x,2) for x in d.get("exchange_wait_us_per_query",[])]))
except Exception as e:
    print("N=$N failed", e)
PY
  tail -3 gpurun_out/${TAG}_bench_n$N.err
done
