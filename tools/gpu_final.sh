#!/usr/bin/env bash
# Final measurements of a round: full GPU test suite, bench lines (batch64, utt10s, T sweep), ncu launch list, ncu --set full of
# the dominant kernel, reference arms, and (last, because it can take the GPU down) the two-stream pair-kernel diagnosis.
set -u
tag="${1:-final}"; shift || true
steps="${*:-tests bench launches full utt sweep hang}"
O=gpurun_out
mkdir -p $O
has() { case " $steps " in *" $1 "*) return 0;; *) return 1;; esac; }
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*"; }
if has tests; then
  lap "pytest -m gpu (whole suite)"
  timeout 2400 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/tests_${tag}.log 2>&1
  grep -E "passed|failed|error" $O/tests_${tag}.log | tail -3; grep -E "^(FAILED|ERROR)" $O/tests_${tag}.log | head -20
fi
if has bench; then
  lap "bench default (batch64)"
  timeout 900 python bench.py > $O/bench_${tag}.json 2> $O/bench_${tag}.err; tail -c 2500 $O/bench_${tag}.json
fi
if has utt; then
  lap "bench utt10s"
  timeout 600 python bench.py --workload utt10s --steps 5 > $O/bench_utt10s_${tag}.json 2> $O/bench_utt10s_${tag}.err; tail -c 600 $O/bench_utt10s_${tag}.json
  timeout 300 python tools/stage_times.py utt10s 100 fast > $O/stage_utt10s_${tag}.json 2>$O/stage_utt10s_${tag}.err; cat $O/stage_utt10s_${tag}.json
  timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_${tag}.json 2> $O/stage_${tag}.err; cat $O/stage_${tag}.json
fi
if has launches; then
  lap "ncu launch list, one batch64 step"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_batch64_${tag}.csv python tools/profile_step.py batch64 100 > $O/launches_${tag}.log 2>&1
  echo "rc=$?"; python tools/summarize_launches.py $O/launches_batch64_${tag}.csv "ncu launch list, batch64 T=100 (${tag})" > $O/launches_${tag}.md; head -30 $O/launches_${tag}.md
fi
if has full; then
  lap "ncu --set full: mel denoiser GEMMs"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_tc2 -c 6 -f -o $O/prof_mel_${tag} python tools/profile_mel.py batch64 2 > $O/prof_mel_${tag}.log 2>&1
  echo "rc=$?"; ls -la $O/prof_mel_${tag}.ncu-rep
fi
if has sweep; then
  lap "bench --workload sweep (configs[4])"
  timeout 900 python bench.py --workload sweep --steps 2 --warmup 1 > $O/bench_sweep_${tag}.json 2> $O/bench_sweep_${tag}.err; tail -c 400 $O/bench_sweep_${tag}.json
fi
if has hang; then
  lap "two F0 nets forked at batch64 with the pair guard OFF (round-1 hang configuration), no dual schedule"
  SSB_TC_NO_DUAL=1 SSB_F0_FORK_ALWAYS=1 SSB_TC_PAIR_CONCURRENT=1 timeout -s KILL 120 python tools/stage_times.py batch64 20 fast > $O/fork_concurrent_${tag}.json 2> $O/fork_concurrent_${tag}.err; echo "rc=$?"; cat $O/fork_concurrent_${tag}.json; tail -2 $O/fork_concurrent_${tag}.err
  nvidia-smi --query-gpu=name,utilization.gpu,memory.used --format=csv,noheader | head -1
fi
lap "done"
