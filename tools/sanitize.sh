#!/bin/bash
# compute-sanitizer passes over the kernel unit tests (the kernels carry hand-rolled mbarrier / TMA / TMEM
# protocols). Run on a GPU box:  bash tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
# Logs: gpurun_out/sanitize_<tool>.log; the summary line of each tool is printed at the end.
TOOL=${1:-memcheck}
KEXPR=${2:-"gemm or layernorm or attention or ddpm or vq or misc"}
mkdir -p gpurun_out
export FDM_B200_SANITIZE=1   # tests shrink their largest shapes under the sanitizer
timeout ${SANITIZE_TIMEOUT:-900} compute-sanitizer --tool "$TOOL" --target-processes all --error-exitcode 9 \
  --log-file gpurun_out/sanitize_${TOOL}.log \
  python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$KEXPR" > gpurun_out/sanitize_${TOOL}_pytest.log 2>&1
rc=$?
echo "compute-sanitizer --tool $TOOL: exit code $rc"
tail -3 gpurun_out/sanitize_${TOOL}_pytest.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error" gpurun_out/sanitize_${TOOL}.log | sort | uniq -c | sort -rn | head -20
exit $rc
