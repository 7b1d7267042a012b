#!/usr/bin/env bash
# Which of TMA+MMA / epilogue bounds the gate and the residual GEMM of a mel-denoiser layer at batch64 scale?
# One sampler step under ncu per setting of SSB_TC_DEBUG (1: epilogue only drains TMEM, 6: no TMA loads and no MMAs),
# one launch per GEMM (no dual schedule), general pair kernel (the tap-reuse kernel has no probe bits).
O=gpurun_out; mkdir -p $O
for dbg in 0 1 6; do
  SSB_TC_NO_DUAL=1 SSB_TC_NO_TAP_REUSE=1 SSB_TC_DEBUG=$dbg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none \
    --profile-from-start off -k regex:conv_gemm_tc2_kernel --csv --log-file $O/probe_layer_dbg${dbg}.csv python tools/profile_mel.py batch64 1 > $O/probe_layer_dbg${dbg}.log 2>&1
  python tools/summarize_launches.py $O/probe_layer_dbg${dbg}.csv "SSB_TC_DEBUG=$dbg" 2>/dev/null | grep -E "conv_gemm" | head -6
done
