#!/bin/bash
# time every exp_so/libwfe_*.so variant (and the shipped library): tools/time_variants.sh [kinds...]
cd "$(dirname "$0")/.."
kinds=${@:-noise}
for so in exp_so/libwfe_*.so; do
  for k in $kinds; do
    WFE_LIB_OVERRIDE=$PWD/$so timeout 120 python tools/time_kernel.py 256 128 10 $k 2>&1 | tail -1
  done
done
