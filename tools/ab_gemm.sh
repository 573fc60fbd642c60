#!/bin/bash
# A/B on ONE box: round-1 GEMM kernel (tools/_trace/libfdm_r1gemm.so) vs HEAD, per-shape and per-step
mkdir -p gpurun_out
for i in 1 2; do
  echo "== r1 gemm, pass $i"; FDM_B200_LIB=$PWD/tools/_trace/libfdm_r1gemm.so python tools/gemm_shapes.py 2>&1 | grep "M="
  echo "== HEAD gemm, pass $i"; python tools/gemm_shapes.py 2>&1 | grep "M="
done
for i in 1 2; do
  echo "== r1 gemm bench $i"; FDM_B200_LIB=$PWD/tools/_trace/libfdm_r1gemm.so python bench.py --ddpm-steps 200 --steps 2 --warmup 1 --named none --microbench none --skip-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_denoise_step'], d['roofline']['frac'], d['clocks'])"
  echo "== HEAD bench $i"; python bench.py --ddpm-steps 200 --steps 2 --warmup 1 --named none --microbench none --skip-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_denoise_step'], d['roofline']['frac'], d['clocks'])"
done
