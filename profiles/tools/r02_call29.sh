#!/bin/bash
# Final build (gausslet parabasal loops unrolled 3 / 2): A/B against 6 / 2, full GPU suite, smoke, ncu --set full of the
# shipped gausslet and plain-ray k_shade, default bench as the driver runs it, small-workload bench lines.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
{
for w in config5_1e6; do for l in librpx.so librpx_m62.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > $O/r02_c29_ab_${w}_${l%.so}.log 2>&1
  tail -1 $O/r02_c29_ab_${w}_${l%.so}.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], '%.3f'%d['roofline']['frac'])" || echo "$w $l FAILED"
done; done
} > $O/r02_c29_ab.log 2>&1
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/r02_c29_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_c29_smoke.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 3 -f -o $O/prof_r02g_gauss_final \
    python bench.py --workload config5_1e6 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_c29_ncu_gauss.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 2 -f -o $O/prof_r02g_config2_final \
    python bench.py --workload config2 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_c29_ncu_config2.log 2>&1
(time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5) > $O/r02_c29_bench_default.log 2>&1
for w in config5_1e6 config2; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_c29_bench_$w.log 2>&1
done
cat $O/r02_c29_ab.log; tail -4 $O/r02_c29_tests.log; tail -3 $O/r02_c29_smoke.log; tail -c 600 $O/r02_c29_bench_default.log
