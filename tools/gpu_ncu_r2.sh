#!/bin/bash
# ncu evidence of the streamed single-query path (round 2); outputs under gpurun_out/
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== full-size index: duration + DRAM bytes of gather_solo, single pass, no cache flush"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
  -k regex:gather_solo -s 40 -c 12 --csv --log-file gpurun_out/r2_gather_dram_fullsize.csv python tools/stream_probe.py --queries 32 "" > gpurun_out/r2_ncu1.log 2>&1
tail -3 gpurun_out/r2_ncu1.log; tail -4 gpurun_out/r2_gather_dram_fullsize.csv | cut -c1-300
echo "== launch list of bench.py"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --prewarm 0 --no-cpu-baseline > gpurun_out/r2_ncu2.log 2>&1
tail -2 gpurun_out/r2_ncu2.log | cut -c1-400; wc -l gpurun_out/r2_launches_bench.csv
echo "== set full: gather_solo + reduce_kernel on the m = 2.5 M twin"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gather_solo|reduce_kernel' -s 30 -c 3 -f -o gpurun_out/r2_gather_full \
  python tools/stream_probe.py --m 2500000 --queries 16 "" > gpurun_out/r2_ncu3.log 2>&1
tail -3 gpurun_out/r2_ncu3.log; ls -la gpurun_out/*.ncu-rep
