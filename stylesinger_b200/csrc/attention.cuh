#pragma once
#include "common.cuh"

namespace ssb {

struct AttnArgs {
  const int4* utt_q = nullptr;  // query layout table
  const int4* utt_k = nullptr;  // key/value layout table
  int B = 0;
  int max_q = 0;                // max query length (grid sizing)
  int heads = 2;                // head dim is fixed at 128
  const float* Q = nullptr; int ldq = 0;
  const float* K = nullptr; int ldk = 0;
  const float* V = nullptr; int ldv = 0;
  const float* keymask = nullptr;  // per key row (guarded), 1 = attend, 0 = -inf; may be null
  float scale = 1.0f;              // applied to q (hd^-0.5)
  float* out = nullptr; int ldo = 0;
};

int attention(Ctx& ctx, const AttnArgs& a);

}  // namespace ssb
