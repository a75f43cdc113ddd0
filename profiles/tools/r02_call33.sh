#!/bin/bash
# Final round-2 build (lean child staging for every kernel, parabasal loops unrolled 3 / 2): full GPU suite, smoke,
# default bench as the driver runs it, ncu --set full of the gausslet k_shade, bench lines of the small workloads.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/r02_c33_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_c33_smoke.log 2>&1
(time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5) > $O/r02_c33_bench_default.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 3 -f -o $O/prof_r02h_gauss_final \
    python bench.py --workload config5_1e6 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_c33_ncu_gauss.log 2>&1
for w in config5_1e6 config2 config4_prisms config5_rays config3 mesh; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_c33_bench_$w.log 2>&1
done
tail -4 $O/r02_c33_tests.log; tail -3 $O/r02_c33_smoke.log
for w in default config5_1e6 config2 config4_prisms config5_rays config3 mesh; do grep "^{" $O/r02_c33_bench_$w.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], '%.3f'%d['roofline']['frac'], d['roofline']['per_launch_ms'])"; done
