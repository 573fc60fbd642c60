#!/bin/bash
# A/B of an alternative library build against the in-tree one on ONE box, alternating:  bash tools/ab_lib.sh <lib.so> [bench args...]
ALT=$1; shift
for i in 1 2; do
  for lib in "$ALT" ""; do
    echo "== ${lib:-HEAD} pass $i: $*"
    FDM_B200_LIB=${lib:+$PWD/$lib} python bench.py --ddpm-steps 200 --steps 2 --warmup 1 --named none --microbench none --skip-cpu-baseline "$@" 2>&1 | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_denoise_step'], d['roofline']['frac'], d['clocks']['sm_mhz'])"
  done
done
