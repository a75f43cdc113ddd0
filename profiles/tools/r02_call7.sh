#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_consume_gpu.py tests/test_capture.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r02_c7_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c7_smoke.log 2>&1
bash profiles/tools/r02_divergence.sh > gpurun_out/r02_c7_div.log 2>&1
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_shade" -c 18 --csv --log-file gpurun_out/r02_traffic_consume.csv \
    python bench.py --rays 4000000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_c7_traffic.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --rays 4000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
(time timeout 900 python bench.py --steps 10 --warmup 3) > gpurun_out/r02_c7_bench_full.log 2>&1
tail -6 gpurun_out/r02_c7_tests.log; tail -4 gpurun_out/r02_c7_smoke.log; tail -c 700 gpurun_out/r02_c7_bench_full.log
