#!/bin/bash
# compute-sanitizer passes over the single-GPU kernel tests (run on a B200 box; slow).
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck]
tool=${1:-memcheck}
cd "$(dirname "$0")/.."
exec compute-sanitizer --tool "$tool" --error-exitcode 1 --launch-timeout 120 \
  python -m pytest tests/test_dft_gemm_gpu.py -x -q -k "rowmajor or scatter"
