#!/usr/bin/env bash
# Which of TMA / MMA / epilogue / L2 prefetch bounds the gate and the residual GEMM of a mel-denoiser layer at batch64 scale?
# One sampler step under ncu per setting, one launch per GEMM (no dual schedule), general pair kernel (the tap-reuse kernel
# has no probe bits).  SSB_TC_DEBUG bits: 1 = epilogue only drains TMEM, 2 = no MMAs, 4 = no TMA loads.
# usage: tools/probe_layer.sh ["dbg[:ENV=VAL]" ...]     default: 0 1 0:SSB_TC_NO_L2_PREFETCH=1 1:SSB_TC_NO_L2_PREFETCH=1 2 3
O=gpurun_out; mkdir -p $O
cfgs="${*:-0 1 0:SSB_TC_NO_L2_PREFETCH=1 1:SSB_TC_NO_L2_PREFETCH=1 2 3}"
for cfg in $cfgs; do
  dbg="${cfg%%:*}"; extra=""; [ "$cfg" != "$dbg" ] && extra="${cfg#*:}"
  tag="dbg${dbg}${extra:+_${extra%%=*}}"
  env SSB_TC_NO_DUAL=1 SSB_TC_NO_TAP_REUSE=1 SSB_TC_DEBUG=$dbg $extra timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
    --profile-from-start off -k regex:conv_gemm_tc2_kernel --csv --log-file $O/probe_layer_${tag}.csv python tools/profile_mel.py batch64 1 > $O/probe_layer_${tag}.log 2>&1
  echo "-- SSB_TC_DEBUG=$dbg $extra"
  python tools/summarize_launches.py $O/probe_layer_${tag}.csv "$tag" 2>/dev/null | grep -E "conv_gemm" | head -3
done
