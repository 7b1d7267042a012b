#!/usr/bin/env bash
# Round-2 visit 4: fast gate epilogue + dual kernel + attention_tc + tiled noise conv: tests, A/B stage times, ncu.
set -u
tag="${1:-v4}"; shift || true
steps="${*:-tests ab full launches}"
O=gpurun_out
mkdir -p $O
has() { case " $steps " in *" $1 "*) return 0;; *) return 1;; esac; }
t0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - t0 )) s] $*"; }
if has tests; then
  lap "attention_tc op tests first (new kernel: a trap here must not hide the rest)"
  timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -rA -p no:cacheprovider -k "attention_tc_op" > $O/tests_attn_${tag}.log 2>&1
  grep -E "passed|failed|error" $O/tests_attn_${tag}.log | tail -2; grep -E "^(FAILED|ERROR)|L-inf|rc=" $O/tests_attn_${tag}.log | head -12
  lap "targeted tests"
  timeout 1800 python -m pytest tests/test_gpu_scale.py tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -rA -p no:cacheprovider -k "not attention_tc_op" > $O/tests_${tag}.log 2>&1
  grep -E "passed|failed|error" $O/tests_${tag}.log | tail -3
  grep -E "^(FAILED|ERROR)|L-inf|agree|max .diff" $O/tests_${tag}.log | head -40
fi
if has ab; then
  lap "stage times (default)"
  timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_${tag}.json 2> $O/stage_${tag}.err; cat $O/stage_${tag}.json
  lap "stage times (SSB_TC_NO_DUAL=1)"
  SSB_TC_NO_DUAL=1 timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_nodual_${tag}.json 2> $O/stage_nodual_${tag}.err; cat $O/stage_nodual_${tag}.json
  lap "stage times (SSB_ATTN_TC=1)"
  SSB_ATTN_TC=1 timeout 200 python tools/stage_times.py batch64 100 fast > $O/stage_attn_${tag}.json 2> $O/stage_attn_${tag}.err; cat $O/stage_attn_${tag}.json; tail -2 $O/stage_attn_${tag}.err
fi
if has full; then
  lap "ncu --set full: mel denoiser GEMMs"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm_tc2 -c 6 -f -o $O/prof_mel_${tag} python tools/profile_mel.py batch64 2 > $O/prof_mel_${tag}.log 2>&1
  echo "rc=$?"; ls -la $O/prof_mel_${tag}.ncu-rep
fi
if has launches; then
  lap "ncu launch list, one batch64 step"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_batch64_${tag}.csv python tools/profile_step.py batch64 100 > $O/launches_${tag}.log 2>&1
  echo "rc=$?"; python tools/summarize_launches.py $O/launches_batch64_${tag}.csv "ncu launch list, batch64 T=100 (${tag})" > $O/launches_${tag}.md; head -34 $O/launches_${tag}.md
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
