#!/bin/sh
# times every variants/*.so on the given configs: sh tools/time_variants.sh "1 2"
for so in variants/*.so; do
  echo "### $so"
  LUCID_B200_SO=$PWD/$so timeout 300 python tools/gpu_probe.py $1 2>&1 | grep -E "stage_ms"
done
