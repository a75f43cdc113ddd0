#!/bin/bash
# Final round-2 pass on one B200: full GPU suite, smoke, ncu launch list + k_shade DRAM traffic of the default
# bench command, the default bench exactly as the driver runs it, the reference arm, and the other workloads.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_g_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_g_smoke.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_g_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_g_launches_bench.log 2>&1
timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"k_shade" -c 18 --csv \
    --log-file gpurun_out/r02_g_traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_g_traffic.log 2>&1
(time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r02_g_bench_default.log 2>&1
(time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/r02_g_bench_reference.log 2>&1
(time timeout 600 python bench.py --workload config4 --steps 3 --warmup 1 --no-cpu-baseline) > gpurun_out/r02_g_bench_config4.log 2>&1
for w in config2 config3 config4_prisms config5_rays config5_1e6 mesh; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_g_bench_$w.log 2>&1
done
tail -8 gpurun_out/r02_g_tests.log; tail -3 gpurun_out/r02_g_smoke.log
for f in default reference config4; do echo "== $f"; tail -c 500 gpurun_out/r02_g_bench_$f.log; done
