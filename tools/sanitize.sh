#!/bin/bash
# compute-sanitizer passes over the sm_100a kernels (run on a B200 box; slow -- the tests below use small shapes).
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck]        1 GPU: GEMM, fused pointwise, FFT, engine step
#   N=2 tools/sanitize.sh memcheck                                     + the peer-scatter / barrier path on 2 GPUs
# Logs go to gpurun_out/sanitize_<tool>[_2gpu].log; the exit code is the sanitizer's.
tool=${1:-memcheck}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS="compute-sanitizer --tool $tool --error-exitcode 1 --launch-timeout 120 --target-processes all"
rc=0
$CS python -m pytest tests/test_dft_gemm_gpu.py tests/test_fused_pointwise_gpu.py tests/test_fft_radix_gpu.py \
    tests/test_spectral_in_gpu.py -x -q -k "rowmajor or scatter or spectral_out or dpre_dw or head or forward_inverse or spectral_in" \
    > gpurun_out/sanitize_${tool}.log 2>&1 || rc=$?
tail -n 5 gpurun_out/sanitize_${tool}.log
if [ "${N:-1}" -ge 2 ]; then
  DFNO_TEST_WORLD=2 $CS python -m pytest tests/test_fused_multigpu.py tests/test_p2p_multigpu.py -x -q \
      > gpurun_out/sanitize_${tool}_2gpu.log 2>&1 || rc=$?
  tail -n 5 gpurun_out/sanitize_${tool}_2gpu.log
fi
exit $rc
