#!/bin/bash
# One multi-GPU box session (round 2): correctness at the full world size, both bench arms, the comm
# measurements and BASELINE configs 3 / 4 -- everything time-boxed, everything logged under gpurun_out/.
#   gpurun --gpus 8 --timeout 1200 -- 'tools/multi_gpu_session.sh'
#   STAGES="tests bench ref" tools/multi_gpu_session.sh      # subset
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
N=$(python -c "import torch; print(torch.cuda.device_count())")
PORT=29700
STAGES=${STAGES:-"tests bench ref exposed a2a profile cfg4 cfg4ref cfg3 cfg3ref"}
tr() { PORT=$((PORT + 1)); echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"; }
run() { local t=$1 log=$2; shift 2; echo "[$(date +%T)] $* -> $log"; timeout "$t" "$@" > "$OUT/$log" 2> "$OUT/$log.err"; echo "    rc=$? $(tail -c 300 "$OUT/$log" | tr '\n' ' ')"; }
case $N in
  8) P3="1 1 2 2 2 1"; P4="1 1 1 1 1 8" ;;
  4) P3="1 1 2 2 1 1"; P4="1 1 1 1 1 4" ;;
  *) P3="1 1 2 1 1 1"; P4="1 1 1 1 1 2" ;;
esac
for s in $STAGES; do
  case $s in
    tests)   run 600 "multigpu_tests_${N}gpu.log" python -m pytest tests/test_fused_multigpu.py tests/test_p2p_multigpu.py -q -s --tb=short ;;
    bench)   run 240 "bench_fused_${N}gpu.json" $(tr) bench.py --gpus $N --steps 20 --warmup 5 ;;
    ref)     run 400 "bench_reference_${N}gpu.json" $(tr) bench.py --gpus $N --steps 6 --warmup 3 --impl reference ;;
    base)    run 300 "bench_baseline_${N}gpu.json" $(tr) bench.py --gpus $N --steps 4 --warmup 3 --impl baseline ;;
    exposed) run 200 "exposed_a2a_${N}gpu.log" $(tr) benchmarks/exposed_a2a.py ;;
    a2a)     run 240 "a2a_sweep_${N}gpu.log" $(tr) benchmarks/a2a_sweep.py ;;
    profile) run 200 "profile_step_${N}gpu.txt" $(tr) benchmarks/profile_step.py --out $OUT/profile_step_${N}gpu_table.txt ;;
    cfg4)    run 240 "cfg4_fused_${N}gpu.json" $(tr) bench.py --gpus $N --steps 10 --warmup 3 --grid 64 --nt 32 --width 24 --modes 8 8 8 8 --tin 8 --partition $P4 ;;
    cfg4ref) run 300 "cfg4_reference_${N}gpu.json" $(tr) bench.py --gpus $N --steps 4 --warmup 3 --grid 64 --nt 32 --width 24 --modes 8 8 8 8 --tin 8 --partition $P4 --impl reference ;;
    cfg3)    run 300 "cfg3_fused_${N}gpu.json" $(tr) bench.py --gpus $N --steps 5 --warmup 3 --grid 256 --nt 16 --width 32 --modes 12 12 12 8 --in-channels 2 --partition $P3 ;;
    weak)    # weak-scaling series (benchmarks/gen_scripts.py --system b200): every point with <= N ranks
             (cd benchmarks && python gen_scripts.py --system b200 --clean-old > /dev/null
              for n in 1 2 4 8; do [ $n -le $N ] || continue
                for f in grad_weak_scaling_spatial_gpu grad_weak_scaling_temporal_gpu; do timeout 200 ./$f.sh $n > ../$OUT/weak_${f}_$n.log 2>&1; done
              done
              python - <<'PY' > ../gpurun_out/weak_scaling_summary.txt
import glob, json, os
rows = []
for d in ("grad_weak_scaling_spatial_gpu", "grad_weak_scaling_temporal_gpu"):
    for f in sorted(glob.glob(os.path.join(d, "*-0-*.json"))):
        j = json.load(open(f)); n = int(f.rsplit("-", 1)[1].split(".")[0])
        rows.append((d, n, os.path.basename(f), j.get("dt"), j.get("dt_grad"), j.get("dt_comm")))
for r in sorted(rows):
    print(f"{r[0]:34s} ranks {r[1]}  fwd {r[3]*1e3:8.3f} ms  bwd {r[4]*1e3:8.3f} ms  comm {r[5]*1e3:7.3f} ms   {r[2]}")
PY
             ); cat $OUT/weak_scaling_summary.txt ;;
    cfg3ref) run 400 "cfg3_reference_${N}gpu.json" $(tr) bench.py --gpus $N --steps 3 --warmup 3 --grid 256 --nt 16 --width 32 --modes 12 12 12 8 --in-channels 2 --partition $P3 --impl reference ;;
  esac
done
ls -la $OUT | tail -n 30
