#!/usr/bin/env bash
# Round-2 visit 3: interleaved dual kernel (tc2d) + tap-reuse kernel: targeted tests, A/B stage times, ncu launch list + full captures.
set -u
tag="${1:-v3}"; shift || true
steps="${*:-tests ab launches full utt}"
O=gpurun_out
mkdir -p $O
has() { case " $steps " in *" $1 "*) return 0;; *) return 1;; esac; }
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*"; }
if has tests; then
  lap "targeted tests"
  timeout 1500 python -m pytest tests/test_gpu_scale.py tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider > $O/tests_${tag}.log 2>&1
  grep -E "passed|failed|error" $O/tests_${tag}.log | tail -3
  grep -E "^(FAILED|ERROR)|L-inf|agree" $O/tests_${tag}.log | head -30
fi
if has ab; then
  lap "stage times (default: dual + tap reuse)"
  timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_${tag}.json 2> $O/stage_${tag}.err; cat $O/stage_${tag}.json
  lap "stage times (SSB_TC_NO_DUAL=1)"
  SSB_TC_NO_DUAL=1 timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_nodual_${tag}.json 2> $O/stage_nodual_${tag}.err; cat $O/stage_nodual_${tag}.json
  lap "stage times (SSB_TC_NO_DUAL=1 SSB_TC_NO_TAP_REUSE=1)"
  SSB_TC_NO_DUAL=1 SSB_TC_NO_TAP_REUSE=1 timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_nodual_noreuse_${tag}.json 2> $O/stage_nodual_noreuse_${tag}.err; cat $O/stage_nodual_noreuse_${tag}.json
fi
if has launches; then
  lap "ncu launch list, one batch64 step"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_batch64_${tag}.csv python tools/profile_step.py batch64 100 > $O/launches_${tag}.log 2>&1
  echo "rc=$?"; python tools/summarize_launches.py $O/launches_batch64_${tag}.csv "ncu launch list, batch64 T=100 (${tag})" | head -30
fi
if has full; then
  lap "ncu --set full: mel denoiser GEMMs (hoist, in_proj, gate, dual x4, ...)"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_tc2 -c 8 -f -o $O/prof_mel_${tag} python tools/profile_mel.py batch64 2 > $O/prof_mel_${tag}.log 2>&1
  echo "rc=$?"; ls -la $O/prof_mel_${tag}.ncu-rep
fi
if has utt; then
  lap "utt10s stage times"
  timeout 300 python tools/stage_times.py utt10s 100 fast > $O/stage_utt10s_${tag}.json 2>$O/stage_utt10s_${tag}.err; cat $O/stage_utt10s_${tag}.json
fi
if has bench; then
  lap "bench default"
  timeout 900 python bench.py > $O/bench_${tag}.json 2> $O/bench_${tag}.err; tail -c 1500 $O/bench_${tag}.json
fi
lap "done"
