#!/usr/bin/env bash
# Round-2 visit 2: targeted tests of the new kernels, A/B of the tap-reuse kernel, ncu launch list + full captures, fork diagnosis.
set -u
tag="${1:-v2}"; shift || true
steps="${*:-tests ab launches full bench hang}"
O=gpurun_out
mkdir -p $O
has() { case " $steps " in *" $1 "*) return 0;; *) return 1;; esac; }
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*"; }
if has tests; then
  lap "targeted tests"
  timeout 1200 python -m pytest tests/test_gpu_scale.py tests/test_gpu_reference_dropin.py tests/test_gpu_dataset_tool.py tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider > $O/tests_${tag}.log 2>&1
  grep -E "passed|failed|error" $O/tests_${tag}.log | tail -3
  grep -E "^(FAILED|ERROR)" $O/tests_${tag}.log | head -20
fi
if has ab; then
  lap "stage times (tap reuse on)"
  timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_${tag}.json 2> $O/stage_${tag}.err; cat $O/stage_${tag}.json
  lap "stage times (SSB_TC_NO_TAP_REUSE=1)"
  SSB_TC_NO_TAP_REUSE=1 timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_noreuse_${tag}.json 2> $O/stage_noreuse_${tag}.err; cat $O/stage_noreuse_${tag}.json
fi
if has launches; then
  lap "ncu launch list, one batch64 step"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_batch64_${tag}.csv python tools/profile_step.py batch64 100 > $O/launches_${tag}.log 2>&1
  echo "rc=$?"; python tools/summarize_launches.py $O/launches_batch64_${tag}.csv "ncu launch list, batch64 T=100 (${tag})" | head -24
fi
if has full; then
  lap "ncu --set full: hoist GEMM, in_proj, 2 x (GATE, RES_SKIP) of the mel denoiser"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_tc2 -c 6 -f -o $O/prof_mel_${tag} python tools/profile_mel.py batch64 2 > $O/prof_mel_${tag}.log 2>&1
  echo "rc=$?"; ls -la $O/prof_mel_${tag}.ncu-rep
fi
if has bench; then
  lap "bench default"
  timeout 600 python bench.py > $O/bench_${tag}.json 2> $O/bench_${tag}.err; tail -c 1200 $O/bench_${tag}.json
fi
if has hang; then
  lap "forward with the two F0 nets forked at batch64, pair guard ON"
  SSB_F0_FORK_ALWAYS=1 timeout -s KILL 150 python tools/stage_times.py batch64 20 fast > $O/fork_guard_${tag}.json 2> $O/fork_guard_${tag}.err; echo "rc=$?"; cat $O/fork_guard_${tag}.json
  lap "same, pair guard OFF (the round-1 hang configuration)"
  SSB_F0_FORK_ALWAYS=1 SSB_TC_PAIR_CONCURRENT=1 timeout -s KILL 150 python tools/stage_times.py batch64 20 fast > $O/fork_concurrent_${tag}.json 2> $O/fork_concurrent_${tag}.err; echo "rc=$?"; cat $O/fork_concurrent_${tag}.json; tail -2 $O/fork_concurrent_${tag}.err
  nvidia-smi --query-gpu=name,utilization.gpu,memory.used --format=csv,noheader | head -1
fi
lap "done"
