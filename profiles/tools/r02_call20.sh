#!/bin/bash
# Re-entry validation of the round-2 build on one B200: full GPU suite, pooled page-locked result arrays of the
# drop-in (e2e_dropin of a short consume-mode run), bench lines of the mesh workloads with the final kernels,
# 256-ray tiles at 2 CTAs / SM, ld.global.cg parent reads and lean child staging at 4 CTAs / SM (more L1) against the
# shipped build (full JSON lines kept), ncu --set full (with
# source) of the shipped gausslet and plain-ray k_shade.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
O=gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > $O/r02_c20_tests.log 2>&1
(time timeout 600 python bench.py --rays 8000000 --steps 3 --warmup 3 --no-cpu-baseline) > $O/r02_c20_bench_consume8e6.log 2>&1
for w in mesh mesh_large; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_c20_bench_$w.log 2>&1
done
{
for rep in 1; do for w in config5_1e6 config2; do for l in librpx_old.so librpx.so librpx_t256.so librpx_ldcg.so librpx_lean4.so; do
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline \
      > $O/r02_c20_ab_${w}_${l%.so}_$rep.log 2>&1
  tail -1 $O/r02_c20_ab_${w}_${l%.so}_$rep.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'], '%.3f'%d['roofline']['frac'])" || echo "$w $l FAILED"
done; done; done
for l in librpx_t256.so librpx_ldcg.so librpx_lean4.so; do
  echo "parity under $l"
  RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -x -q -k "config5 or config2 or config4 or zoo" 2>&1 | tail -2
done
} > $O/r02_c20_ab.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 3 -f -o $O/prof_r02b_gauss \
    python bench.py --workload config5_1e6 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_c20_ncu_gauss.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade -c 2 -f -o $O/prof_r02b_config2 \
    python bench.py --workload config2 --steps 1 --warmup 0 --no-cpu-baseline > $O/r02_c20_ncu_config2.log 2>&1
cat $O/r02_c20_tests.log $O/r02_c20_ab.log
tail -c 1800 $O/r02_c20_bench_consume8e6.log
for w in mesh mesh_large; do tail -c 700 $O/r02_c20_bench_$w.log; echo; done
ls -la $O | tail -24
