#!/usr/bin/env bash
# Round-2 GPU visit: one gpurun call that collects everything that needs the box (see the step list).
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash tools/gpu_visit.sh <tag> [steps...]'
# steps (default: all): tests stage bench group sweep arms smoke_ncu hang
set -u
tag="${1:-v1}"; shift || true
steps="${*:-tests stage bench group sweep arms smoke_ncu hang}"
O=gpurun_out
mkdir -p $O
has() { case " $steps " in *" $1 "*) return 0;; *) return 1;; esac; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader | head -1
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*"; }

if has tests; then
  lap "pytest -m gpu"
  timeout 1200 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/tests_${tag}.log 2>&1
  grep -E "passed|failed|error" $O/tests_${tag}.log | tail -3
  grep -E "^(FAILED|ERROR)" $O/tests_${tag}.log | head -20
fi
if has stage; then
  lap "stage times batch64"
  timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_${tag}.json 2> $O/stage_${tag}.err; cat $O/stage_${tag}.json
fi
if has bench; then
  lap "bench default"
  timeout 600 python bench.py > $O/bench_${tag}.json 2> $O/bench_${tag}.err; tail -c 1500 $O/bench_${tag}.json; tail -3 $O/bench_${tag}.err
  lap "bench utt10s"
  timeout 200 python bench.py --workload utt10s --steps 5 --no-cpu-baseline > $O/bench_utt10s_${tag}.json 2> $O/bench_utt10s_${tag}.err; tail -c 600 $O/bench_utt10s_${tag}.json
fi
if has group; then
  for gf in 9472 18944 37888; do
    lap "stage times, utterance groups of $gf frames"
    SSB_MEL_GROUP_FRAMES=$gf SSB_F0_GROUP_FRAMES=$gf timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_group${gf}_${tag}.json 2> $O/stage_group${gf}_${tag}.err
    cat $O/stage_group${gf}_${tag}.json
  done
fi
if has sweep; then
  lap "T sweep (configs[4])"
  timeout 900 python bench.py --workload sweep --steps 2 > $O/sweep_${tag}.json 2> $O/sweep_${tag}.err; tail -c 2500 $O/sweep_${tag}.json; tail -3 $O/sweep_${tag}.err
fi
if has arms; then
  lap "reference arms (BASELINE.md section 3)"
  timeout 900 python tools/baseline_arms.py --out $O/baseline_arms_${tag}.json > $O/baseline_arms_${tag}.log 2>&1; tail -60 $O/baseline_arms_${tag}.json
fi
if has smoke_ncu; then
  lap "smoke under ncu (launch census)"
  env | grep -i -E "nsight|compute_profiler|injection" | head
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_smoke_${tag}.csv \
      python -c "import os; print({k: v for k, v in os.environ.items() if any(s in k.upper() for s in ('NSIGHT', 'PROFILER', 'INJECTION'))}); import __graft_entry__ as g; g.smoke()" > $O/smoke_ncu_${tag}.log 2>&1
  echo "ncu rc=$?"; tail -5 $O/smoke_ncu_${tag}.log; grep -c "conv_gemm_tc" $O/launches_smoke_${tag}.csv
fi
if has hang; then
  lap "two-stream reproducer, guard on (default)"
  timeout -s KILL 120 python tools/repro_two_stream_hang.py 10 > $O/hang_guard_${tag}.log 2>&1; echo "rc=$?"; tail -2 $O/hang_guard_${tag}.log
  lap "two-stream reproducer, guard off (SSB_TC_PAIR_CONCURRENT=1)"
  SSB_TC_PAIR_CONCURRENT=1 timeout -s KILL 90 python tools/repro_two_stream_hang.py 10 > $O/hang_concurrent_${tag}.log 2>&1; echo "rc=$?"; tail -2 $O/hang_concurrent_${tag}.log
  nvidia-smi --query-gpu=name,utilization.gpu,memory.used --format=csv,noheader | head -1
fi
lap "done"
