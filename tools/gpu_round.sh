#!/bin/bash
# One gpurun call of the development loop: smoke, targeted parity tests, bench (+ timeline), full GPU suite.
# usage: tools/gpu_round.sh <tag> [pytest -k expression for the targeted pass]
TAG=${1:-dev}
KEXPR=${2:-"exchange or solo_path or zero_copy or device_pointer or search_sequence or golden or config1"}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== targeted tests" ; timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_targeted.log
echo "== bench" ; timeout 900 python bench.py --steps 20 --warmup 5 --timeline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -40 gpurun_out/${TAG}_bench.err
echo "== full gpu suite" ; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_full.log
