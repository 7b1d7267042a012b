// Implicit-GEMM 1-D convolution over guard-banded ragged rows (see common.cuh).
//
//   Y[r, n] = epilogue( sum_{j<taps} sum_{c<Cin} f(A[r + (j - center)*dil, c]) * W[j][c][n] )
//
// covers every dense operator of the hot path: Linear (taps=1), Conv1d k=3/5/7/9/11 with dilation,
// and ConvTranspose1d(k=2u, stride u) as a 3-tap conv whose N dimension is (phase, Cout).
// Rows outside an utterance are zero in A (guard bands) => zero "same" padding for free.
#pragma once
#include "common.cuh"

namespace ssb {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_LRELU = 3, ACT_TANH = 4, ACT_MISH = 5 };
enum EpiMode { EPI_GENERIC = 0, EPI_GATE = 1, EPI_RES_SKIP = 2 };

struct Epi {
  int mode = EPI_GENERIC;
  const float* bias = nullptr;   // [N]
  const float* add = nullptr;    // [rows, ld_add]  v += add[r, n]  (before alpha/act)
  int ld_add = 0;
  float alpha = 1.0f;            // v *= alpha
  int act = ACT_NONE;
  float act_slope = 0.1f;
  const float* res = nullptr;    // v = (v + res[r, n]) * beta
  int ld_res = 0;
  float beta = 1.0f;
  const float* rowmask = nullptr;  // v *= rowmask[r]
  float* out = nullptr;
  int ldo = 0;
  int accum = 0;                 // out = (out + v) * gamma
  float gamma = 1.0f;
  float* out2 = nullptr;         // out2[r, n] = v + vec2[n]
  int ldo2 = 0;
  const float* vec2 = nullptr;
  __half* out2_h = nullptr;      // optional: (v + vec2) additionally split into fp16 hi/lo planes [rows, ldh]
  __half* out2_l = nullptr;      // (A operand of the tensor-core GEMM that consumes it)
  int ldh = 0;
  int plane_act = ACT_NONE;      // activation applied to the plane value (ACT_LRELU: pre-activation of the consumer conv)
  float plane_slope = 0.1f;
  // EPI_RES_SKIP: columns [0,C) -> residual path (res/beta/rowmask/out/out2), [C,2C) -> skip
  float* skip = nullptr;
  int ld_skip = 0;
  int C = 0;
  int skip_init = 0;             // 1: skip = v, 0: skip += v
};

struct ConvGemm {
  const float* A = nullptr;
  int lda = 0;
  int Cin = 0;       // multiple of 16
  int taps = 1;
  int dil = 1;
  int center = 0;    // tap j reads row r + (j - center) * dil
  const float* W = nullptr;  // [taps][Cin][Npad], Npad multiple of 4
  int N = 0;
  int Npad = 0;
  const int2* tiles = nullptr;
  int ntiles = 0;
  int a_act = ACT_NONE;   // ACT_NONE or ACT_LRELU applied to A on load
  float a_slope = 0.1f;
  float a_scale = 1.0f;   // A multiplied by this before the activation
  Epi e;
};

// Enqueue on ctx.stream (no-op in dry mode). Returns 0 or a negative error code.
int conv_gemm(Ctx& ctx, const ConvGemm& p);

}  // namespace ssb
