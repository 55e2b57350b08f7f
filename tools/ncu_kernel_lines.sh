#!/bin/bash
# Source-level profile of one kernel of one config: bash tools/ncu_kernel_lines.sh <tag> <config> <kernel> [top]
# -> gpurun_out/<tag>_<kernel>_config<c>_lines.txt and _functions.txt (the report itself is dropped)
tag=$1; c=$2; k=$3; top=${4:-60}
out=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name $k --launch-skip 1 -c 1 \
  -o $out/tmp_$k -f python tools/profile_frame.py $c 2 > $out/${tag}_${k}_config${c}.log 2>&1
ncu -i $out/tmp_$k.ncu-rep --page source --csv --print-source cuda,sass > $out/tmp_$k.csv 2>/dev/null
python tools/ncu_lines.py $out/tmp_$k.csv $top > $out/${tag}_${k}_config${c}_lines.txt 2>&1
python tools/ncu_funcs.py $out/tmp_$k.csv > $out/${tag}_${k}_config${c}_functions.txt 2>&1
rm -f $out/tmp_$k.ncu-rep $out/tmp_$k.csv
