#!/bin/bash
# One GPU-box session, staged and time-boxed, everything logged under gpurun_out/ (the only directory
# that comes back from a gpurun call).  Usage (from the repo root, on the box):
#
#   tools/gpu_session.sh tests bench1 launches          # pick stages; DRY=1 prints the commands only
#   N=8 tools/gpu_session.sh benchN exposed a2a         # multi-GPU stages use N ranks (default: all GPUs)
#
# Stages:
#   tests      pytest -m gpu (includes the non-strict experimental file last)           ~3 min
#   smoke      __graft_entry__.smoke()                                                   ~20 s
#   bench1     bench.py, 1 GPU, fused and baseline arms                                  ~2 min
#   benchN     bench.py on $N GPUs, fused (+ DFNO_STAGED_SCATTER=0/1 A/B) and baseline   ~3 min
#   refarm     bench.py --impl reference (unmodified reference on baseline/compat) on 1 and $N GPUs   ~3 min
#   mgtests    pytest tests/test_fused_multigpu.py tests/test_p2p_multigpu.py on all GPUs (log kept)   ~2 min
#   exposed    benchmarks/exposed_a2a.py on $N GPUs (direct vs staged, same process)     ~40 s
#   a2a        benchmarks/a2a_sweep.py on $N GPUs                                        ~1 min
#   launches   ncu launch list of one training step (1 GPU)                              ~1 min
#   ncu        ncu --set full of the top kernels (1 GPU; never a multi-rank command)     ~4 min
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
OUTABS="$PWD/$OUT"
NGPU=$(python - <<'PY'
import torch
print(torch.cuda.device_count())
PY
)
N=${N:-$NGPU}
PORT=${PORT:-29610}
run() {   # run <seconds> <logfile> <command...>
  local t=$1 log=$2; shift 2
  echo "[$(date +%T)] $* (limit ${t}s) -> $OUT/$log"
  if [ "${DRY:-0}" != 0 ]; then return 0; fi
  timeout "$t" "$@" > "$OUT/$log" 2>&1
  local rc=$?
  echo "    rc=$rc  $(tail -n 1 "$OUT/$log" | cut -c1-200)"
  return 0
}
trun() {  # trun <nproc>: sets $TR to a torchrun prefix with a fresh rendezvous port
  PORT=$((PORT + 1))
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $PORT"
}
for stage in "$@"; do
  case "$stage" in
    tests)    run 600 tests.log python -m pytest tests -q -m gpu -rxXs ;;
    smoke)    run 120 smoke.log python -c "import __graft_entry__ as g; g.smoke()" ;;
    bench1)   run 240 bench_fused_1gpu.json python bench.py --gpus 1 --steps 10 --warmup 3
              run 300 bench_baseline_1gpu.json python bench.py --gpus 1 --steps 4 --warmup 3 --impl baseline ;;
    benchN)   trun "$N"; run 240 "bench_fused_${N}gpu.json" $TR bench.py --gpus "$N" --steps 20 --warmup 5
              for s in 0 1 r3; do
                trun "$N"; DFNO_STAGED_SCATTER=$s run 240 "bench_fused_${N}gpu_staged$s.json" $TR bench.py --gpus "$N" --steps 20 --warmup 5 --no-e2e
              done
              trun "$N"; run 300 "bench_baseline_${N}gpu.json" $TR bench.py --gpus "$N" --steps 6 --warmup 3 --impl baseline ;;
    refarm)   run 400 bench_reference_1gpu.json python bench.py --gpus 1 --steps 4 --warmup 3 --impl reference
              trun "$N"; run 400 "bench_reference_${N}gpu.json" $TR bench.py --gpus "$N" --steps 6 --warmup 3 --impl reference ;;
    mgtests)  run 900 "multigpu_tests_${N}gpu.log" python -m pytest tests/test_fused_multigpu.py tests/test_p2p_multigpu.py -q -s -rxXs ;;
    exposed)  trun "$N"; run 120 "exposed_a2a_${N}gpu.log" $TR benchmarks/exposed_a2a.py ;;
    a2a)      trun "$N"; run 180 "a2a_sweep_${N}gpu.log" $TR benchmarks/a2a_sweep.py ;;
    launches) run 240 launches.log ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
                  --log-file "$OUT/launch_list.csv" python benchmarks/one_step.py ;;
    ncu)      for k in dft_gemm head_bwd2 head_fwd spectral_out dpre_dw mix_ lift_bwd; do
                run 240 "ncu_$k.log" ncu --set full --clock-control none --import-source on -k "regex:$k" -c 3 -f \
                    -o "$OUT/ncu_$k" python benchmarks/one_step.py
              done ;;
    *) echo "unknown stage $stage"; exit 2 ;;
  esac
done
ls -la "$OUT" | tail -n 30
