// fp32 SIMT implicit-GEMM conv1d (see conv_gemm.cuh).  128 x TN output tile per CTA, 256 threads,
// 8 x (TN/16) register tile per thread, K chunks of 16 double-buffered through shared memory with
// register prefetch.  All global loads are 16-byte vectors along the contiguous (channel / N) axis.
#include "conv_gemm.cuh"

namespace ssb {

namespace {

constexpr int TM = TILE_M;
constexpr int TK = 16;
constexpr int AS_LD = TM + 4;

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_GELU: return gelu_erf(v);
    case ACT_LRELU: return v > 0.0f ? v : v * slope;
    case ACT_TANH: return tanhf(v);
    case ACT_MISH: return mishf_(v);
    default: return v;
  }
}

// Epilogue for 4 consecutive output columns n..n+3 of row r.
__device__ __forceinline__ void epilogue4(const Epi& e, int64_t r, int n, int N, float4 a4) {
  float v[4] = {a4.x, a4.y, a4.z, a4.w};
  if (e.mode == EPI_GATE) {
    // columns are packed (sigmoid-arg, tanh-arg) pairs; N is the packed width (2*C)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      int na = n + 2 * q;
      if (na + 1 < N) {
        float g = v[2 * q] + (e.bias ? e.bias[na] : 0.0f);
        float f = v[2 * q + 1] + (e.bias ? e.bias[na + 1] : 0.0f);
        if (e.add) {
          g += e.add[r * e.ld_add + na];
          f += e.add[r * e.ld_add + na + 1];
        }
        float z = sigmoidf_(g) * tanhf(f);
        if (e.rowmask) z *= e.rowmask[r];
        e.out[r * e.ldo + (na >> 1)] = z;
      }
    }
    return;
  }
  if (e.mode == EPI_RES_SKIP) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int nn = n + q;
      if (nn >= N) continue;
      float x = v[q] + (e.bias ? e.bias[nn] : 0.0f);
      if (nn < e.C) {
        x = (x + e.res[r * e.ld_res + nn]) * e.beta;
        if (e.rowmask) x *= e.rowmask[r];
        e.out[r * e.ldo + nn] = x;
        if (e.out2) e.out2[r * e.ldo2 + nn] = x + e.vec2[nn];
      } else {
        int s = nn - e.C;
        float* sp = e.skip + r * e.ld_skip + s;
        *sp = e.skip_init ? x : (*sp + x);
      }
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int nn = n + q;
    if (nn >= N) continue;
    float x = v[q];
    if (e.bias) x += e.bias[nn];
    if (e.add) x += e.add[r * e.ld_add + nn];
    x *= e.alpha;
    x = apply_act(x, e.act, e.act_slope);
    if (e.res) x = (x + e.res[r * e.ld_res + nn]) * e.beta;
    if (e.rowmask) x *= e.rowmask[r];
    if (e.accum) x = (x + e.out[r * e.ldo + nn]) * e.gamma;
    e.out[r * e.ldo + nn] = x;
    if (e.out2) e.out2[r * e.ldo2 + nn] = x + (e.vec2 ? e.vec2[nn] : 0.0f);
    if (e.out2_h) {
      float y = x + (e.vec2 ? e.vec2[nn] : 0.0f);
      if (e.plane_act == ACT_LRELU) y = y > 0.0f ? y : y * e.plane_slope;
      const __half h = __float2half_rn(y);
      e.out2_h[r * e.ldh + nn] = h;
      e.out2_l[r * e.ldh + nn] = __float2half_rn(y - __half2float(h));
    }
  }
}

template <int TN>
__global__ void __launch_bounds__(256, 2) conv_gemm_kernel(const ConvGemm p) {
  constexpr int NG = TN / 64;            // column groups of 64 (each thread: 4 cols per group)
  constexpr int BV = TK * TN / 4 / 256;  // float4 B loads per thread (2 for TN=128, 1 for TN=64)
  __shared__ __align__(16) float As[2][TK][AS_LD];
  __shared__ __align__(16) float Bs[2][TK][TN];

  const int2 tile = p.tiles[blockIdx.x];
  const int64_t row0 = tile.x;
  const int nvalid = tile.y;
  const int n0 = blockIdx.y * TN;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;

  const int kchunks = p.Cin / TK;
  const int nk = p.taps * kchunks;

  float acc[8][4 * NG];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4 * NG; ++j) acc[i][j] = 0.0f;

  float4 ra[2];
  float4 rb[BV];

  auto gload = [&](int it) {
    const int tap = it / kchunks;
    const int c0 = (it - tap * kchunks) * TK;
    const int64_t shift = (int64_t)(tap - p.center) * p.dil;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      const int row = idx >> 2, kq = (idx & 3) << 2;
      const float* src = p.A + (row0 + row + shift) * (int64_t)p.lda + c0 + kq;
      float4 v = __ldg(reinterpret_cast<const float4*>(src));
      if (p.a_scale != 1.0f) { v.x *= p.a_scale; v.y *= p.a_scale; v.z *= p.a_scale; v.w *= p.a_scale; }
      if (p.a_act == ACT_LRELU) {
        v.x = v.x > 0.f ? v.x : v.x * p.a_slope;
        v.y = v.y > 0.f ? v.y : v.y * p.a_slope;
        v.z = v.z > 0.f ? v.z : v.z * p.a_slope;
        v.w = v.w > 0.f ? v.w : v.w * p.a_slope;
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int idx = tid + i * 256;
      const int k = idx / (TN / 4), n4 = (idx % (TN / 4)) * 4;
      const int n = n0 + n4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < p.Npad) v = __ldg(reinterpret_cast<const float4*>(p.W + ((int64_t)(tap * p.Cin + c0 + k)) * p.Npad + n));
      rb[i] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * 256;
      const int row = idx >> 2, kq = (idx & 3) << 2;
      As[buf][kq + 0][row] = ra[i].x;
      As[buf][kq + 1][row] = ra[i].y;
      As[buf][kq + 2][row] = ra[i].z;
      As[buf][kq + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int idx = tid + i * 256;
      const int k = idx / (TN / 4), n4 = (idx % (TN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][k][n4]) = rb[i];
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();

  for (int it = 0; it < nk; ++it) {
    const int buf = it & 1;
    if (it + 1 < nk) gload(it + 1);
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[8], b[4 * NG];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
        b[4 * g + 0] = b0.x; b[4 * g + 1] = b0.y; b[4 * g + 2] = b0.z; b[4 * g + 3] = b0.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4 * NG; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (it + 1 < nk) sstore(buf ^ 1);
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
    if (rl >= nvalid) continue;
    const int64_t r = row0 + rl;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int n = n0 + g * 64 + tx * 4;
      if (n >= p.N) continue;
      epilogue4(p.e, r, n, p.N, make_float4(acc[i][4 * g + 0], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]));
    }
  }
}

}  // namespace

int conv_gemm(Ctx& ctx, const ConvGemm& p) {
  SSB_CHECK(p.Cin % TK == 0, "conv_gemm: Cin must be a multiple of 16");
  SSB_CHECK(p.Npad % 4 == 0 && p.Npad >= p.N, "conv_gemm: bad Npad");
  SSB_CHECK(p.lda % 4 == 0, "conv_gemm: lda must be a multiple of 4");
  SSB_CHECK(p.e.out != nullptr || ctx.dry, "conv_gemm: no output");
  if (ctx.dry || p.ntiles == 0) return 0;
  SSB_CHECK((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0,
            "conv_gemm: A/W must be 16-byte aligned");
  if (p.N > 64) {
    dim3 grid(p.ntiles, (p.N + 127) / 128);
    conv_gemm_kernel<128><<<grid, 256, 0, ctx.stream>>>(p);
  } else {
    dim3 grid(p.ntiles, 1);
    conv_gemm_kernel<64><<<grid, 256, 0, ctx.stream>>>(p);
  }
  SSB_CUDA(cudaGetLastError());
  ++g_launches;
  return 0;
}

}  // namespace ssb
