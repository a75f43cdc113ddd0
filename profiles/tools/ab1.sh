#!/bin/bash
# one timed bench per (lib, workload), each under its own timeout: profiles/tools/ab1.sh "lib1.so lib2.so" "config2 config5"
for w in $2; do for l in $1; do
RPX_LIB=$PWD/raypier_optics_b200/csrc/$l timeout 180 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w $l', '%.4g'%d['value'], '%.4f'%d['ms_per_step'], d['roofline']['per_launch_ms'])" || echo "$w $l FAILED"
done; done
