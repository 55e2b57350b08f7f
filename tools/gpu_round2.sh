#!/bin/bash
# One measurement pass of round 2 on a B200 box (run under gpurun): parity suite, smoke, the default bench line
# (configs[3], 4K), the reference arm, bench lines of the other configs, stage probe, the ncu launch list of the
# default bench command and --set full captures of one frame.  Everything lands in gpurun_out/.
#   bash tools/gpu_round2.sh <tag> [what: "tests bench ncu" default all] [full-capture configs, default "3 1"]
tag=${1:-r2}
what=${2:-"tests bench others ncu"}
full=${3:-"3 1"}
lpf=${LAUNCHES_PER_FRAME:-11}   # kernels per frame incl. k_tie_runs and k_info_out
out=gpurun_out
mkdir -p $out
nproc > $out/${tag}_nproc.txt
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1

if [[ $what == *tests* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $out/${tag}_pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
fi
if [[ $what == *bench* ]]; then
  timeout 900 python bench.py > $out/${tag}_bench_config3.json 2> $out/${tag}_bench_config3.err
fi
if [[ $what == *others* ]]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>&1
  for c in 1 2; do
    timeout 900 python bench.py --config $c --steps 10 > $out/${tag}_bench_config$c.json 2> $out/${tag}_bench_config$c.err
  done
  timeout 900 python tools/gpu_probe.py 0 1 2 3 > $out/${tag}_probe_stage_times.txt 2>&1
fi
if [[ $what == *ncu* ]]; then
  # launch list of the bench command itself (per-launch times are cold-cache and serialised)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file $out/${tag}_launches_config3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sustained-seconds 0.01 \
    > $out/${tag}_launches_bench.log 2>&1
  # full captures: second frame of profile_frame.py (frame 1 is warm-up: k_pad_positions + $lpf launches per frame)
  for c in $full; do
    timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip $((lpf + 1)) -c $lpf \
      -o $out/${tag}_full_config$c -f python tools/profile_frame.py $c 2 > $out/${tag}_full_config$c.log 2>&1
    python tools/ncu_summary.py $out/${tag}_full_config$c.ncu-rep > $out/${tag}_ncu_full_config$c.txt 2>&1
    traffic="$traffic config$c=$out/${tag}_full_config$c.ncu-rep"
  done
  python tools/ncu_traffic.py $traffic > $out/${tag}_dram_traffic.json 2>$out/${tag}_dram_traffic.err
  # per-function instruction / stall-sample shares of the shading kernel, then drop the reports
  # (gpurun only copies 64 MiB back)
  for c in $full; do
    for k in ${NCU_KERNELS:-k_block_shade k_block_sort}; do
      ncu -i $out/${tag}_full_config$c.ncu-rep --page source --csv --print-source cuda,sass --kernel-name $k > $out/src_$c.csv 2>/dev/null
      python tools/ncu_funcs.py $out/src_$c.csv > $out/${tag}_${k}_config${c}_functions.txt 2>&1
    done
    rm -f $out/src_$c.csv $out/${tag}_full_config$c.ncu-rep
  done
fi
rm -f $out/config*.png
ls -la $out | tail -40
