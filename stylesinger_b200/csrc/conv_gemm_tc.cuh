// tcgen05 / TMA implicit-GEMM conv1d for the dense contractions of the denoisers.
//
// Precision: every fp32 operand is carried as TWO fp16 planes (hi = fp16(x), lo = fp16(x - hi));
// each K step issues three tcgen05.mma (hi*hi + hi*lo + lo*hi) into one fp32 TMEM accumulator, so the
// contraction keeps ~22 mantissa bits (single-pass bf16/fp16/TF32 cannot meet the
// mel L-inf < 1e-3 bar at T=100; a 3-pass split can).  Effective tensor peak = 1/3 of the fp16 peak.
//
// Layout: A = activation planes [rows, C] fp16 row-major (guard-banded rows, see common.cuh), loaded by
// TMA as [128 rows x 64 ch] boxes with 128B swizzle at row offset (tap - center) * dilation;
// B = weights [taps*N, Cin] fp16 (K-major), boxes [BN n x 64 ch] (single-CTA kernel) or [hb x 64] (CTA-pair kernel:
// each CTA of a 2-CTA cluster stages half of a 2*hb-wide N tile and its own 128 rows of A; the leader issues
// cta_group::2 MMAs with M = 256).  D = fp32 in TMEM, double buffered so the epilogue of tile i overlaps the MMAs
// of tile i+1 (persistent CTAs).  The pair kernel is used when there are >= num_SMs pair tiles, the 64-wide
// single-CTA kernel otherwise (conv_gemm_tc()).
// An optional SECOND operand pair (A2, W2: 1 tap, no shift) extends the K loop: the DiffNet layer
// GEMM contracts [3 taps x C of y | 256 of cond] in one accumulator, so the conditioner projection
// needs neither a hoisted [rows, L*2C] fp32 buffer nor an epilogue read.
// Env switches (diagnostics): SSB_TC_NO_PAIR=1 disables the pair kernel, SSB_TC_PAIR_CONCURRENT=1 lifts the cross-stream
// serialisation of pair kernels (conv_gemm_tc.cu), SSB_TC_BN256=1 enables 256-wide single-CTA
// tiles, SSB_TC_DEBUG=<bits> switches parts of the kernel off for tools/gemm_probe.py (results are then garbage).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "conv_gemm.cuh"

namespace ssb {

struct ConvTC {            // packed weights for the tensor-core path
  __half* W_hi = nullptr;  // [taps][N][Cin]
  __half* W_lo = nullptr;
  CUtensorMap tm_hi[3], tm_lo[3];  // [0]: box 128 rows (BN=128), [1]: box 64 rows (BN=64), [2]: box 256 rows (BN=256)
  CUtensorMap tm2_hi, tm2_lo;       // CTA-pair kernel: box hb rows (each CTA of a pair stages half of a 2*hb-wide N tile)
  int hb = 0;                       // 0: no pair packing
  int taps = 1, Cin = 0, N = 0, dil = 1, center = 0;
  const float* bias = nullptr;  // [N] (packed column order)
  bool ok = false;
};

struct EpiTC {
  int mode = EPI_GENERIC;        // EPI_GENERIC: out = acc + bias ; EPI_GATE ; EPI_RES_SKIP
  const float* bias = nullptr;
  float* out = nullptr;          // GENERIC / RES_SKIP residual path: fp32 [rows, ldo]
  int ldo = 0;
  __half* oh = nullptr;          // GATE: z planes [rows, C];  RES_SKIP: y = x_new + vec2 planes [rows, C] (may be null)
  __half* ol = nullptr;
  int ldh = 0;
  const float* add = nullptr;    // GATE: pre-activation addend [rows, ld_add] in packed (sigmoid, tanh) column order - the
  int ld_add = 0;                //   hoisted, step-invariant conditioner projection of this layer (stages.cu prepare_cond)
  const float* res = nullptr;    // RES_SKIP: x [rows, ld_res]
  int ld_res = 0;
  const __half* rh = nullptr;    // RES_SKIP, planes-only residual stream: x = (rh + rl) - vec1 is read from the fp16 hi/lo
  const __half* rl = nullptr;    //   planes of y = x + step bias (the GATE GEMM's own A operand); no fp32 x is read or written
  int ld_rh = 0;
  const float* vec1 = nullptr;   //   step bias of the CURRENT layer [C]
  float beta = 1.0f;
  const float* vec2 = nullptr;   // RES_SKIP: step bias of the next layer [C]
  float* skip = nullptr;         // RES_SKIP: [rows, ld_skip]
  int ld_skip = 0;
  int C = 0;
  int skip_init = 0;
  // EPI_GENERIC extras: v = act(acc + bias); v += res; out = accum ? (out + v) * gamma : v; planes = plane_act(v)
  int act = ACT_NONE;
  float act_slope = 0.1f;
  int accum = 0;
  float gamma = 1.0f;
  int plane_act = ACT_NONE;      // activation applied to the value written to the fp16 planes (pre-activation of the consumer)
  float plane_slope = 0.1f;
  float alpha = 1.0f;            // GENERIC: v = act((acc + bias) * alpha)   (ACT_GELU supported here for the FFT FFN)
  const float* rowmask = nullptr;  // GENERIC: v = (v + res) * rowmask[row]
  int n_valid = 0;               // > 0: output columns >= n_valid are padding (weights padded to a tile multiple): skipped
  int skip_tiled = 0;            // RES_SKIP: skip accumulator stored chunk-tiled - [row tile][32-row quarter][32-col chunk][32 rows][32 cols]
                                 //   fp32, so that every 32 x 32 epilogue chunk is one contiguous 4 KB block (private to this epilogue)
  int tile_base = 0;             //   index of tiles[0] in the full tile table (a GEMM over a sub-range of the row tiles)
  int out_nb = 0;                // GENERIC, > 0: `out` is column-block-major: block j = columns [j*out_nb, (j+1)*out_nb) is its own
  int64_t out_bs = 0;            //   [rows, out_nb] matrix at out + j * out_bs (the hoisted conditioner: one matrix per layer)
  int stream_hints = 1;          // skip accumulator and conditioner addends are touched once per launch: ld/st.global.cs (evict-first)
                                 //   so that they do not push the y / z planes (re-read by the next launch) out of L2; SSB_TC_NO_STREAM_HINTS=1: off
  int l2_prefetch = 1;           // warp 3 pulls the next tile's epilogue operands into L2 (SSB_TC_NO_L2_PREFETCH=1: off)
  __half* sh = nullptr;          // RES_SKIP (last layer): the finished skip sum also as fp16 planes [rows, C]
  __half* sl = nullptr;
};

struct GemmTC {
  const __half* A_hi = nullptr;  // [rows_total, Cin]
  const __half* A_lo = nullptr;
  int64_t rows_total = 0;
  const ConvTC* w = nullptr;
  const __half* A2_hi = nullptr; // optional second operand [rows_total, w2->Cin], contracted with w2 (1 tap)
  const __half* A2_lo = nullptr;
  const ConvTC* w2 = nullptr;
  const int2* tiles = nullptr;
  int ntiles = 0;
  EpiTC e;
};

bool tc_available();  // driver entry point for cuTensorMapEncodeTiled resolved
int make_weight_maps(ConvTC* w);
// TMA descriptor of an activation plane [rows, cols] fp16 (box 128 rows x 64 cols, 128B swizzle)
int make_act_map(CUtensorMap* m, const void* ptr, int64_t rows, int cols, int box_rows = 128);
int conv_gemm_tc(Ctx& ctx, const GemmTC& p);
// One launch for two INDEPENDENT problems, interleaved tile by tile: g = a 3-tap gate conv (EPI_GATE), r = a 1x1 residual conv
// (EPI_RES_SKIP).  Falls back to two launches when the shapes do not qualify.  Opt-in: ssb_set_interleaved_layers(1) / SSB_TC_DUAL=1.
int conv_gemm_tc_dual(Ctx& ctx, const GemmTC& g, const GemmTC& r);
bool dual_enabled();
int set_dual_enabled(int on);  // process-wide switch (default off; SSB_TC_DUAL=1 starts with it on)
// diagnostics: launches of one kernel variant ("tc2<128,GATE>", "tc<64,GENERIC>", ...), the variants seen so far,
// and the activation-descriptor cache counters
long long variant_launch_count(const char* name);
int variant_names(char* buf, int cap);
void tensor_map_cache_stats(long long* encodes, long long* hits);
// x fp32 [rows, ld] -> hi/lo planes [rows, C] (all rows incl. guards; guards stay zero)
int split_planes(Ctx& ctx, const float* x, int ld, int64_t rows, int C, float scale, __half* hi, __half* lo);

}  // namespace ssb
