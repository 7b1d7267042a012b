#!/usr/bin/env bash
# One GPU-box visit that answers "is the tree still good, and how fast is it": meant to be the single command of a
# gpurun call (every call costs ~1.5 GPU-minutes of box set-up on top of the run time, so batch the steps).
#
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_check.sh [quick|full] [tag]'
#
# quick: tcgen05 tests + stage times + short bench  (~3 min)
# full : all GPU tests + stage times + default bench (CPU baseline, latency) (~6 min)
# Results land in gpurun_out/ (merged back by gpurun): tests_<tag>.log, stage_<tag>.json, bench_<tag>.json
set -u
mode="${1:-quick}"
tag="${2:-$(date +%H%M%S)}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader | head -1
if [ "$mode" = "full" ]; then
  timeout 500 python -m pytest tests -m gpu -x -q > "gpurun_out/tests_${tag}.log" 2>&1
else
  timeout 300 python -m pytest tests/test_gpu_tc.py -x -q > "gpurun_out/tests_${tag}.log" 2>&1
fi
tail -3 "gpurun_out/tests_${tag}.log"
timeout 150 python tools/stage_times.py batch64 100 fast > "gpurun_out/stage_${tag}.json" 2> "gpurun_out/stage_${tag}.err"
cat "gpurun_out/stage_${tag}.json"
tail -2 "gpurun_out/stage_${tag}.err"
if [ "$mode" = "full" ]; then
  timeout 300 python bench.py > "gpurun_out/bench_${tag}.json" 2> "gpurun_out/bench_${tag}.err"
else
  timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-latency > "gpurun_out/bench_${tag}.json" 2> "gpurun_out/bench_${tag}.err"
fi
python - "gpurun_out/bench_${tag}.json" <<'EOF'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print("bench: %.0f frames/s, %.0f ms/step, e2e %.0f, roofline %.3f (mel stage %.0f ms), clocks %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], r["frac"], r["stage_ms"], d["clocks"]))
except Exception as e:  # noqa: BLE001
    print("bench line unreadable:", e)
EOF
tail -2 "gpurun_out/bench_${tag}.err"
