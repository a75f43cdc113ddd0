#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CSRC=raypier_optics_b200/csrc
for l in librpx.so librpx_te.so librpx_sf.so; do
  [ -f $CSRC/$l ] || continue
  RPX_LIB=$PWD/$CSRC/$l timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_shade -c 9 --csv --log-file gpurun_out/r02_c9_launch_$l.csv \
    python bench.py --workload config5_1e6 --steps 1 --warmup 0 --no-cpu-baseline > /dev/null 2>&1
  echo "$l: $(grep k_shade gpurun_out/r02_c9_launch_$l.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' ')"
done > gpurun_out/r02_c9_perlaunch.log 2>&1
libs=""; for l in librpx.so librpx_te.so librpx_sf.so librpx_sfte.so; do [ -f $CSRC/$l ] && libs="$libs $l"; done
bash profiles/tools/ab1.sh "$libs $libs" "config5_1e6" > gpurun_out/r02_c9_ab.log 2>&1
cat gpurun_out/r02_c9_perlaunch.log gpurun_out/r02_c9_ab.log
