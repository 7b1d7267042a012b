// tcgen05 + TMA + TMEM implicit-GEMM conv1d (see conv_gemm_tc.cuh).  sm_100a only.
//
// CTA = 256 threads, persistent over (M-tile, N-tile) pairs:
//   warp 0 (1 lane)  TMA producer: per K block loads A_hi, A_lo [128x64] and W_hi, W_lo [BNx64] (128B swizzle)
//   warp 1 (1 lane)  MMA issuer:   4 K-steps x 3 products of tcgen05.mma.kind::f16 (M128 N=BN K16), fp32 in TMEM
//   warp 2           TMEM allocator (2 x BN columns = 2 accumulator buffers)
//   warps 4-7        epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> global,
//                    with the epilogue's own global operands (residual / skip) prefetched one chunk ahead
// smem ring: BN=128: 3 stages x 64 KB, BN=64: 4 stages x 48 KB; mbarriers: full/empty per stage,
// tmem_full/tmem_empty per accumulator buffer.  BN=64 is picked for small problems (more CTAs in flight).
#include <cuda_fp16.h>
#include <stdlib.h>

#include "conv_gemm_tc.cuh"
#include "tc_common.cuh"

namespace ssb {

namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_TILE = BM * BK * 2;  // 16 KB

template <int BN>
struct Cfg {
  static constexpr int B_TILE = BN * BK * 2;
  static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
  static constexpr int STAGES = BN == 256 ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256;
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // BN = 256: the whole 512-column TMEM
};

struct TCParams {
  const int2* tiles;
  int ntiles, NT, taps, kchunks, kchunks2, dil, center, N;
  EpiTC e;
};

using namespace tc;

// Global operands of the epilogue for 32 columns [n, n+32) of row r, fetched ahead of use.
struct Pre {
  float4 a[8];
};
__device__ __forceinline__ void prefetch32(const EpiTC& e, int64_t r, int n, bool valid, Pre& p) {
  if (!valid || (e.n_valid > 0 && n >= e.n_valid)) return;
  if (e.mode == EPI_GENERIC) {
    if (e.res) {
      const float4* rp = reinterpret_cast<const float4*>(e.res + r * e.ld_res + n);
#pragma unroll
      for (int q = 0; q < 8; ++q) p.a[q] = rp[q];
    }
    return;
  }
  if (e.mode != EPI_RES_SKIP) return;
  if (n < e.C) {
    const float4* rp = reinterpret_cast<const float4*>(e.res + r * e.ld_res + n);
#pragma unroll
    for (int q = 0; q < 8; ++q) p.a[q] = rp[q];
  } else if (!e.skip_init) {
    const float4* sp = reinterpret_cast<const float4*>(e.skip + r * e.ld_skip + (n - e.C));
#pragma unroll
    for (int q = 0; q < 8; ++q) p.a[q] = sp[q];
  }
}

// fused epilogue for 32 consecutive columns [n, n+32) of row r
__device__ __forceinline__ void epilogue32(const EpiTC& e, int64_t r, int n, const uint32_t (&raw)[32], const Pre& pre) {
  if (e.n_valid > 0 && n >= e.n_valid) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) + (e.bias ? __ldg(e.bias + n + j) : 0.0f);
  if (e.mode == EPI_GATE) {
    float z[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) z[q] = sigmoidf_(v[2 * q]) * tanhf(v[2 * q + 1]);
    split_store16(e.oh + r * e.ldh + (n >> 1), e.ol + r * e.ldh + (n >> 1), z);
    return;
  }
  if (e.mode == EPI_RES_SKIP) {
    if (n < e.C) {
      float4* op = reinterpret_cast<float4*>(e.out + r * e.ldo + n);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 x0 = pre.a[q];
        v[4 * q] = (v[4 * q] + x0.x) * e.beta;
        v[4 * q + 1] = (v[4 * q + 1] + x0.y) * e.beta;
        v[4 * q + 2] = (v[4 * q + 2] + x0.z) * e.beta;
        v[4 * q + 3] = (v[4 * q + 3] + x0.w) * e.beta;
        op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      if (e.oh) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __ldg(e.vec2 + n + j);
        split_store16(e.oh + r * e.ldh + n, e.ol + r * e.ldh + n, v);
        split_store16(e.oh + r * e.ldh + n + 16, e.ol + r * e.ldh + n + 16, v + 16);
      }
    } else {
      const int sc = n - e.C;
      float4* sp = reinterpret_cast<float4*>(e.skip + r * e.ld_skip + sc);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (!e.skip_init) {
          const float4 o = pre.a[q];
          v[4 * q] += o.x; v[4 * q + 1] += o.y; v[4 * q + 2] += o.z; v[4 * q + 3] += o.w;
        }
        sp[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      if (e.sh) {
        split_store16(e.sh + r * e.C + sc, e.sl + r * e.C + sc, v);
        split_store16(e.sh + r * e.C + sc + 16, e.sl + r * e.C + sc + 16, v + 16);
      }
    }
    return;
  }
  // EPI_GENERIC
  if (e.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (e.act == ACT_LRELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : v[j] * e.act_slope;
  }
  if (e.res) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      v[4 * q] += pre.a[q].x; v[4 * q + 1] += pre.a[q].y; v[4 * q + 2] += pre.a[q].z; v[4 * q + 3] += pre.a[q].w;
    }
  }
  if (e.out) {
    float4* op = reinterpret_cast<float4*>(e.out + r * e.ldo + n);
    if (e.accum) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 o = op[q];
        v[4 * q] = (v[4 * q] + o.x) * e.gamma; v[4 * q + 1] = (v[4 * q + 1] + o.y) * e.gamma;
        v[4 * q + 2] = (v[4 * q + 2] + o.z) * e.gamma; v[4 * q + 3] = (v[4 * q + 3] + o.w) * e.gamma;
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  if (e.oh) {
    if (e.vec2) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += __ldg(e.vec2 + n + j);
    }
    if (e.plane_act == ACT_LRELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.0f ? v[j] : v[j] * e.plane_slope;
    }
    split_store16(e.oh + r * e.ldh + n, e.ol + r * e.ldh + n, v);
    split_store16(e.oh + r * e.ldh + n + 16, e.ol + r * e.ldh + n + 16, v + 16);
  }
}

template <int BN>
__global__ void __launch_bounds__(256, 1)
conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const __grid_constant__ CUtensorMap tmA2_hi, const __grid_constant__ CUtensorMap tmA2_lo,
                    const __grid_constant__ CUtensorMap tmB2_hi, const __grid_constant__ CUtensorMap tmB2_lo,
                    const TCParams p) {
  using K = Cfg<BN>;
  constexpr int STAGES = K::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * K::STAGE);
  // bars: full[STAGES], empty[STAGES], tfull[2], tempty[2]; then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * STAGES, tfull0 = empty0 + 8 * STAGES, tempty0 = tfull0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(K::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int total = p.ntiles * p.NT;
  const int nk1 = p.taps * p.kchunks;
  const int nk = nk1 + p.kchunks2;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int mt = tile / p.NT, nt = tile - mt * p.NT;
        const int row0 = p.tiles[mt].x;
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t fb = full0 + 8 * stage;
          mbar_expect_tx(fb, K::STAGE);
          const uint32_t sa = sbase + stage * K::STAGE;
          if (kb < nk1) {
            const int tap = kb / p.kchunks;
            const int c0 = (kb - tap * p.kchunks) * BK;
            const int arow = row0 + (tap - p.center) * p.dil;
            const int brow = tap * p.N + nt * BN;
            tma_load_2d(sa, &tmA_hi, fb, c0, arow);
            tma_load_2d(sa + A_TILE, &tmA_lo, fb, c0, arow);
            tma_load_2d(sa + 2 * A_TILE, &tmB_hi, fb, c0, brow);
            tma_load_2d(sa + 2 * A_TILE + K::B_TILE, &tmB_lo, fb, c0, brow);
          } else {
            const int c0 = (kb - nk1) * BK;
            tma_load_2d(sa, &tmA2_hi, fb, c0, row0);
            tma_load_2d(sa + A_TILE, &tmA2_lo, fb, c0, row0);
            tma_load_2d(sa + 2 * A_TILE, &tmB2_hi, fb, c0, nt * BN);
            tma_load_2d(sa + 2 * A_TILE + K::B_TILE, &tmB2_lo, fb, c0, nt * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=F32, A=B=F16, both K-major, N=BN, M=128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int a = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(tempty0 + 8 * a, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * BN);
        for (int kb = 0; kb < nk; ++kb) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = sbase + stage * K::STAGE;
          const uint64_t dah = make_sdesc(sa), dal = make_sdesc(sa + A_TILE);
          const uint64_t dbh = make_sdesc(sa + 2 * A_TILE), dbl = make_sdesc(sa + 2 * A_TILE + K::B_TILE);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t off = (uint64_t)((ks * 32) >> 4);  // 16 fp16 = 32 bytes along K inside the swizzle atom
            tc_mma(d_tmem, dah + off, dbh + off, idesc, (kb | ks) != 0 ? 1u : 0u);
            tc_mma(d_tmem, dah + off, dbl + off, idesc, 1u);
            tc_mma(d_tmem, dal + off, dbh + off, idesc, 1u);
          }
          tc_commit(empty0 + 8 * stage);  // smem stage reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull0 + 8 * a);        // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;  // == warp % 4: TMEM lanes [32*ew, 32*ew + 32)
    constexpr int NCH = BN / 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int mt = tile / p.NT, nt = tile - mt * p.NT;
      const int2 t = p.tiles[mt];
      const int rl = ew * 32 + lane;
      const bool valid = rl < t.y;
      const int64_t r = (int64_t)t.x + rl;
      Pre cur, nxt;
      prefetch32(p.e, r, nt * BN, valid, cur);  // issued before the accumulator is ready: overlaps the MMAs
      mbar_wait(tfull0 + 8 * a, aph);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        if (ch + 1 < NCH) prefetch32(p.e, r, nt * BN + (ch + 1) * 32, valid, nxt);
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * BN + ch * 32), v);
        if (valid) epilogue32(p.e, r, nt * BN + ch * 32, v, cur);
        cur = nxt;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * a);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(K::TMEM_COLS) : "memory");
  }
}

__global__ void k_split_planes(const float* x, int ld, int64_t rows, int C, float scale, __half* hi, __half* lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int64_t r = i / C;
  const int c = (int)(i - r * C);
  const float v = x[r * ld + c] * scale;
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_encode_tried = false;

EncodeTiledFn get_encode() {
  if (!g_encode_tried) {
    g_encode_tried = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
  }
  return g_encode;
}

// 2-D fp16 row-major [rows, cols] tensor, box [box_rows x 64 cols], 128B swizzle, zero OOB fill
int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  SSB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult rc = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSB_CHECK(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  return 0;
}

template <int BN>
int launch(Ctx& ctx, const GemmTC& p, const TCParams& tp, int num_sms) {
  const ConvTC& w = *p.w;
  const ConvTC& w2 = p.w2 ? *p.w2 : *p.w;
  const int bi = BN == 128 ? 0 : (BN == 64 ? 1 : 2);
  CUtensorMap ta_hi, ta_lo, ta2_hi, ta2_lo;
  if (make_map(&ta_hi, p.A_hi, (uint64_t)p.rows_total, (uint64_t)w.Cin, BM)) return -1;
  if (make_map(&ta_lo, p.A_lo, (uint64_t)p.rows_total, (uint64_t)w.Cin, BM)) return -1;
  if (p.w2) {
    if (make_map(&ta2_hi, p.A2_hi, (uint64_t)p.rows_total, (uint64_t)w2.Cin, BM)) return -1;
    if (make_map(&ta2_lo, p.A2_lo, (uint64_t)p.rows_total, (uint64_t)w2.Cin, BM)) return -1;
  } else {
    ta2_hi = ta_hi;
    ta2_lo = ta_lo;
  }
  const int total = tp.ntiles * tp.NT;
  const int grid = total < num_sms ? total : num_sms;
  conv_gemm_tc_kernel<BN><<<grid, 256, Cfg<BN>::SMEM, ctx.stream>>>(ta_hi, ta_lo, w.tm_hi[bi], w.tm_lo[bi], ta2_hi, ta2_lo,
                                                                    w2.tm_hi[bi], w2.tm_lo[bi], tp);
  SSB_CUDA(cudaGetLastError());
  ++g_launches;
  return 0;
}

}  // namespace

bool tc_available() { return get_encode() != nullptr; }
int make_act_map(CUtensorMap* m, const void* ptr, int64_t rows, int cols, int box_rows) { return make_map(m, ptr, (uint64_t)rows, (uint64_t)cols, (uint32_t)box_rows); }

int make_weight_maps(ConvTC* w) {
  SSB_CHECK(w->Cin % BK == 0 && w->N % 64 == 0, "tensor-core path needs Cin % 64 == 0 and N % 64 == 0");
  if (w->N % 128 == 0) {
    if (make_map(&w->tm_hi[0], w->W_hi, (uint64_t)w->taps * w->N, (uint64_t)w->Cin, 128)) return -1;
    if (make_map(&w->tm_lo[0], w->W_lo, (uint64_t)w->taps * w->N, (uint64_t)w->Cin, 128)) return -1;
  }
  if (make_map(&w->tm_hi[1], w->W_hi, (uint64_t)w->taps * w->N, (uint64_t)w->Cin, 64)) return -1;
  if (make_map(&w->tm_lo[1], w->W_lo, (uint64_t)w->taps * w->N, (uint64_t)w->Cin, 64)) return -1;
  if (w->N % 256 == 0) {
    if (make_map(&w->tm_hi[2], w->W_hi, (uint64_t)w->taps * w->N, (uint64_t)w->Cin, 256)) return -1;
    if (make_map(&w->tm_lo[2], w->W_lo, (uint64_t)w->taps * w->N, (uint64_t)w->Cin, 256)) return -1;
  }
  w->ok = true;
  return 0;
}

int conv_gemm_tc(Ctx& ctx, const GemmTC& p) {
  if (ctx.dry || p.ntiles == 0) return 0;
  const ConvTC& w = *p.w;
  SSB_CHECK(w.ok && (!p.w2 || (p.w2->ok && p.w2->N == w.N && p.w2->taps == 1)), "conv_gemm_tc: weights not packed for the tensor-core path");
  static bool configured = false;
  static int num_sms = 148;
  if (!configured) {
    SSB_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM));
    SSB_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM));
    SSB_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<256>::SMEM));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    configured = true;
  }
  TCParams tp;
  tp.tiles = p.tiles; tp.ntiles = p.ntiles; tp.taps = w.taps; tp.kchunks = w.Cin / BK;
  tp.kchunks2 = p.w2 ? p.w2->Cin / BK : 0;
  tp.dil = w.dil; tp.center = w.center; tp.N = w.N; tp.e = p.e;
  if (!tp.e.bias) tp.e.bias = w.bias;
  // small problems: 64-wide N tiles keep more SMs busy and shorten each tile's dependent chain
  const bool small = (w.N % 128 != 0) || (int64_t)p.ntiles * (w.N / 128) < (int64_t)num_sms * 2;
  if (small) {
    tp.NT = w.N / 64;
    return launch<64>(ctx, p, tp, num_sms);
  }
  // 256-wide tiles halve the A-operand bytes per FLOP (the kernel is bound by L2->SM operand traffic,
  // profiles/r01_ncu_*) but leave room for only a 2-stage ring (2 x 96 KB) and use the whole TMEM.  Measured on the
  // batch64 mel stage: 1587 ms vs 1441 ms with 128-wide tiles / 3 stages, so it is opt-in (SSB_TC_BN256=1) only.
  static const bool wide_ok = getenv("SSB_TC_BN256") != nullptr;
  if (wide_ok && w.N % 256 == 0 && (!p.w2 || p.w2->N % 256 == 0) && (int64_t)p.ntiles * (w.N / 256) >= (int64_t)num_sms * 2) {
    tp.NT = w.N / 256;
    return launch<256>(ctx, p, tp, num_sms);
  }
  tp.NT = w.N / 128;
  return launch<128>(ctx, p, tp, num_sms);
}

int split_planes(Ctx& ctx, const float* x, int ld, int64_t rows, int C, float scale, __half* hi, __half* lo) {
  if (ctx.dry || rows == 0) return 0;
  const int64_t n = rows * C;
  k_split_planes<<<(unsigned)((n + 255) / 256), 256, 0, ctx.stream>>>(x, ld, rows, C, scale, hi, lo);
  SSB_CUDA(cudaGetLastError());
  ++g_launches;
  return 0;
}

}  // namespace ssb
