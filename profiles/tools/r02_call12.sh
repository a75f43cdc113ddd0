#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02_c12_tests.log 2>&1
bash profiles/tools/ab1.sh "librpx_nc.so librpx.so librpx_nc.so librpx.so" "config4_prisms config2 config5_rays config5_1e6 config4_grating" > gpurun_out/r02_c12_ab.log 2>&1
timeout 600 python bench.py --workload config4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_c12_config4.log 2>&1
RPX_LIB=$PWD/raypier_optics_b200/csrc/librpx_nc.so timeout 600 python bench.py --workload config4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_c12_config4_nc.log 2>&1
tail -8 gpurun_out/r02_c12_tests.log; cat gpurun_out/r02_c12_ab.log
for f in config4 config4_nc; do python - <<PY
import json
for line in open('gpurun_out/r02_c12_$f.log'):
    if line.startswith('{'):
        d=json.loads(line); print('$f', 'value %.4g e2e %.4g frac %.3f'%(d['value'], d['e2e']['value'], d['roofline']['frac']), d['roofline']['per_launch_ms'])
PY
done
