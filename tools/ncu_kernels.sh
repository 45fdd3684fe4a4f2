#!/bin/bash
# One `ncu --set full` pass over the step's own kernels (1 GPU), condensed ON THE BOX into
# gpurun_out/ncu_r2_kernels.json (the .ncu-rep is > 64 MiB and is deleted afterwards).
#   tools/ncu_kernels.sh ['regex'] [count]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
re=${1:-"lift_fwd|mix_fwd|spectral_out|head_fwd|absmax|head_bwd2|dpre_dw|mix_bwd|lift_bwd|adam|permute"}
n=${2:-40}
rep=/tmp/ncu_r2_kernels
timeout 600 ncu --set full --clock-control none -k regex:"$re" -c "$n" -f -o $rep python benchmarks/one_step.py \
    > gpurun_out/ncu_kernels.log 2>&1
tail -n 2 gpurun_out/ncu_kernels.log
python benchmarks/ncu_report.py $rep.ncu-rep --all > gpurun_out/ncu_r2_kernels.json 2> gpurun_out/ncu_report.err
grep -c '"kernel"' gpurun_out/ncu_r2_kernels.json
