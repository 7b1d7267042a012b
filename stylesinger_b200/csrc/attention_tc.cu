// Multi-head attention (2 heads x 128) on tcgen05 tensor cores with TMA-staged tiles.  sm_100a only.
// Rows a3 / a12 of SURVEY.md section 8: self-attention of the FFT blocks (reference modules/commons/common_layers.py:277-286
// -> F.multi_head_attention_forward: q * hd^-0.5, bmm, key-padding -> -inf, fp32 softmax, bmm) and the style aligner's
// cross-attention (modules/StyleSinger/lse.py:41).  Used for long batches; short ones keep the fp32 kernel (attention.cu).
//
// One CTA = 128 queries of one (utterance, head); keys / values stream through a 4-slot shared-memory ring in tiles of 64.
// Both contractions are 3-pass fp16 hi/lo split MMAs like every other tcgen05 GEMM of this library (fp32-class accuracy):
//   S   = Q K^T      M128 x N64  x K128 (head dim),  A = Q planes, B = K planes              -> TMEM, double buffered
//   O  += P V        M128 x N128 x K64  (keys),      A = P (written by the softmax warps),  B = V^T planes -> TMEM
// No running rescale of O: pass 1 streams S once for the exact row maximum m, pass 2 recomputes S, forms
// p = exp(scale * (s - m)) (masked keys -> 0), accumulates l = sum p in registers and O += P V on the tensor cores; the
// epilogue divides by l.  P is carried as fp16 hi/lo planes of 256 * p so that weights down to 2^-33 survive the split.
// The [L, S] score matrix is never materialised (63 MB / utterance / layer in the reference at F = 2812).
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM allocator   warps 4-7: softmax / epilogue (thread = query row)
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "attention.cuh"
#include "conv_gemm_tc.cuh"
#include "tc_common.cuh"

namespace ssb {

namespace {

using namespace tc;

constexpr int AQ = 128, AK = 64, HD = 128;
constexpr int QT = AQ * 64 * 2;            // one [128 x 64] fp16 tile: 16 KB
constexpr int Q_BYTES = 4 * QT;            // hi_h0, hi_h1, lo_h0, lo_h1
constexpr int KT = AK * 64 * 2;            // one [64 keys x 64 d] fp16 tile: 8 KB
constexpr int SLOT = 4 * KT;               // K: hi_h0, hi_h1, lo_h0, lo_h1 ; V^T: hi [128 d x 64 keys], lo
constexpr int NSLOT = 4;
constexpr int PT = AQ * AK * 2;            // one P plane: 16 KB
constexpr int ATT_SMEM = Q_BYTES + NSLOT * SLOT + 2 * PT + 1024 + 256;
constexpr uint32_t ATT_TMEM = 256;         // S: 2 x 64 columns, O: 128 columns

__device__ __forceinline__ void proxy_fence_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct AttnTCParams {
  const int4* utt_q;
  const int4* utt_k;
  int qcol0, kcol0;
  const float* keymask;
  float c2;  // scale * log2(e)
  float* out; int ldo;
  __half* oh; __half* ol; int ldh;
};

__global__ void __launch_bounds__(256, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ_hi, const __grid_constant__ CUtensorMap tmQ_lo,
                    const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
                    const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo, const AttnTCParams p) {
  const int b = blockIdx.z, head = blockIdx.y;
  const int4 uq = p.utt_q[b], uk = p.utt_k[b];
  const int q0 = blockIdx.x * AQ;
  if (q0 >= uq.y) return;  // whole CTA, before any barrier / TMEM state exists
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_BYTES + NSLOT * SLOT + 2 * PT);
  // bars: qfull, kvfull[4], kvempty[4], sfull[2], sempty[2], pready, pfree, ofull ; then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t qbase = smem_u32(smem), ring = qbase + Q_BYTES, pbase = ring + NSLOT * SLOT;
  const uint32_t qfull = smem_u32(bars), kvfull0 = qfull + 8, kvempty0 = kvfull0 + 8 * NSLOT, sfull0 = kvempty0 + 8 * NSLOT;
  const uint32_t sempty0 = sfull0 + 16, pready = sempty0 + 16, pfree = pready + 8, ofull = pfree + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(qfull, 1);
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(kvfull0 + 8 * s, 1);
      mbar_init(kvempty0 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(sfull0 + 8 * a, 1);
      mbar_init(sempty0 + 8 * a, 4);
    }
    mbar_init(pready, 4);
    mbar_init(pfree, 1);
    mbar_init(ofull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(ATT_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  // Key tiles sit on a grid aligned to 8 rows of the K/V layout: the V^T planes are read with the key index as the INNER
  // TMA coordinate, whose byte offset must be a multiple of 16.  The (up to 7) rows in front of the utterance are masked.
  const int klen = uk.y;
  const int kshift = uk.x & 7;
  const int krow0 = uk.x - kshift;
  const int n = (kshift + klen + AK - 1) / AK;  // key tiles
  const int qrow0 = uq.x + q0;
  const int hq = p.qcol0 + head * HD, hk = p.kcol0 + head * HD;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(qfull, Q_BYTES);
      tma_load_2d(qbase, &tmQ_hi, qfull, hq, qrow0);
      tma_load_2d(qbase + QT, &tmQ_hi, qfull, hq + 64, qrow0);
      tma_load_2d(qbase + 2 * QT, &tmQ_lo, qfull, hq, qrow0);
      tma_load_2d(qbase + 3 * QT, &tmQ_lo, qfull, hq + 64, qrow0);
      int slot = 0;
      uint32_t ph = 0;
      auto load_k = [&](int j) {
        mbar_wait(kvempty0 + 8 * slot, ph ^ 1);
        const uint32_t fb = kvfull0 + 8 * slot, sa = ring + slot * SLOT;
        mbar_expect_tx(fb, SLOT);
        tma_load_2d(sa, &tmK_hi, fb, hk, krow0 + j * AK);
        tma_load_2d(sa + KT, &tmK_hi, fb, hk + 64, krow0 + j * AK);
        tma_load_2d(sa + 2 * KT, &tmK_lo, fb, hk, krow0 + j * AK);
        tma_load_2d(sa + 3 * KT, &tmK_lo, fb, hk + 64, krow0 + j * AK);
        if (++slot == NSLOT) { slot = 0; ph ^= 1; }
      };
      auto load_v = [&](int j) {  // V^T planes [heads * 128, key rows]: box = 64 keys x 128 d
        mbar_wait(kvempty0 + 8 * slot, ph ^ 1);
        const uint32_t fb = kvfull0 + 8 * slot, sa = ring + slot * SLOT;
        mbar_expect_tx(fb, SLOT);
        tma_load_2d(sa, &tmV_hi, fb, krow0 + j * AK, head * HD);
        tma_load_2d(sa + 2 * KT, &tmV_lo, fb, krow0 + j * AK, head * HD);
        if (++slot == NSLOT) { slot = 0; ph ^= 1; }
      };
      for (int j = 0; j < n; ++j) load_k(j);  // pass 1: row maxima
      load_k(0);                              // pass 2, in the order the MMA warp consumes: K0, K1, V0, K2, V1, ...
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | ((uint32_t)(AK >> 3) << 17) | ((uint32_t)(AQ >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(AQ >> 4) << 24);
      int slot = 0, g = 0;
      uint32_t ph = 0;
      mbar_wait(qfull, 0);
      tc_fence_after();
      auto issue_s = [&]() {
        const int a = g & 1;
        mbar_wait(sempty0 + 8 * a, ((g >> 1) & 1) ^ 1);
        mbar_wait(kvfull0 + 8 * slot, ph);
        tc_fence_after();
        const uint32_t sa = ring + slot * SLOT;
        const uint32_t d = tmem_base + (uint32_t)(a * AK);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t off = (uint64_t)((ks * 32) >> 4);
            const uint64_t qh = make_sdesc(qbase + h * QT) + off, ql = make_sdesc(qbase + 2 * QT + h * QT) + off;
            const uint64_t kh = make_sdesc(sa + h * KT) + off, kl = make_sdesc(sa + 2 * KT + h * KT) + off;
            tc_mma(d, qh, kh, idesc_s, (h | ks) != 0 ? 1u : 0u);
            tc_mma(d, qh, kl, idesc_s, 1u);
            tc_mma(d, ql, kh, idesc_s, 1u);
          }
        }
        tc_commit(kvempty0 + 8 * slot);
        tc_commit(sfull0 + 8 * a);
        if (++slot == NSLOT) { slot = 0; ph ^= 1; }
        ++g;
      };
      for (int j = 0; j < n; ++j) issue_s();
      issue_s();
      const uint32_t dO = tmem_base + 128u;
      for (int j = 0; j < n; ++j) {
        if (j + 1 < n) issue_s();
        mbar_wait(pready, (uint32_t)(j & 1));
        mbar_wait(kvfull0 + 8 * slot, ph);
        tc_fence_after();
        const uint32_t sv = ring + slot * SLOT;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t off = (uint64_t)((ks * 32) >> 4);
          const uint64_t ph_ = make_sdesc(pbase) + off, pl_ = make_sdesc(pbase + PT) + off;
          const uint64_t vh = make_sdesc(sv) + off, vl = make_sdesc(sv + 2 * KT) + off;
          tc_mma(dO, ph_, vh, idesc_o, (j | ks) != 0 ? 1u : 0u);
          tc_mma(dO, ph_, vl, idesc_o, 1u);
          tc_mma(dO, pl_, vh, idesc_o, 1u);
        }
        tc_commit(kvempty0 + 8 * slot);
        tc_commit(pfree);
        if (++slot == NSLOT) { slot = 0; ph ^= 1; }
      }
      tc_commit(ofull);
    }
  } else if (warp >= 4) {
    const int ew = warp & 3;
    const int row = ew * 32 + lane;  // query row of this thread (= TMEM lane)
    const uint32_t tlane = tmem_base + ((uint32_t)(ew * 32) << 16);
    int g = 0;
    float m = -INFINITY;
    auto key_masks = [&](int j, uint32_t& m0, uint32_t& m1) {
      const int k0 = j * AK + lane - kshift, k1 = k0 + 32;  // key index inside the utterance
      bool v0 = k0 >= 0 && k0 < klen, v1 = k1 >= 0 && k1 < klen;
      if (p.keymask) {
        if (v0) v0 = __ldg(p.keymask + uk.x + k0) != 0.f;
        if (v1) v1 = __ldg(p.keymask + uk.x + k1) != 0.f;
      }
      m0 = __ballot_sync(0xffffffffu, v0);
      m1 = __ballot_sync(0xffffffffu, v1);
    };
    // pass 1: exact row maximum of the raw scores over the valid keys
    for (int j = 0; j < n; ++j, ++g) {
      const int a = g & 1;
      uint32_t m0, m1;
      key_masks(j, m0, m1);
      mbar_wait(sfull0 + 8 * a, (g >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tlane + (uint32_t)(a * AK), v0);
      tmem_ld32(tlane + (uint32_t)(a * AK + 32), v1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty0 + 8 * a);
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if ((m0 >> c) & 1u) m = fmaxf(m, __uint_as_float(v0[c]));
        if ((m1 >> c) & 1u) m = fmaxf(m, __uint_as_float(v1[c]));
      }
    }
    const float mu = m == -INFINITY ? 0.f : m;
    const float c2 = p.c2;
    float l = 0.f;
    const uint32_t prow = pbase + (uint32_t)row * 128u;
    const int sw = row & 7;
    // pass 2: p = exp(scale * (s - m)); O += P V
    for (int j = 0; j < n; ++j, ++g) {
      const int a = g & 1;
      uint32_t m0, m1;
      key_masks(j, m0, m1);
      mbar_wait(sfull0 + 8 * a, (g >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tlane + (uint32_t)(a * AK), v0);
      tmem_ld32(tlane + (uint32_t)(a * AK + 32), v1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sempty0 + 8 * a);
      float pv[64];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        pv[c] = ((m0 >> c) & 1u) ? exp2f((__uint_as_float(v0[c]) - mu) * c2) : 0.f;
        pv[32 + c] = ((m1 >> c) & 1u) ? exp2f((__uint_as_float(v1[c]) - mu) * c2) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 64; ++c) l += pv[c];
      if (j > 0) mbar_wait(pfree, (uint32_t)((j - 1) & 1));  // the MMAs of P_{j-1} V_{j-1} have read the P buffer
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float x0 = pv[ch * 8 + 2 * q] * 256.0f, x1 = pv[ch * 8 + 2 * q + 1] * 256.0f;
          const __half2 hh = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          hw[q] = *reinterpret_cast<const uint32_t*>(&hh);
          lw[q] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const uint32_t dst = prow + (uint32_t)((ch ^ sw) << 4);  // 128B swizzle: 16-byte chunk index XOR (row mod 8)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hw[0]), "r"(hw[1]), "r"(hw[2]), "r"(hw[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + PT), "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3]) : "memory");
      }
      proxy_fence_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(pready);
    }
    // epilogue: O / (256 l)
    mbar_wait(ofull, 0);
    tc_fence_after();
    const float inv = l > 0.f ? 1.0f / (256.0f * l) : 0.f;
    const bool valid = q0 + row < uq.y;
    const int64_t grow = (int64_t)qrow0 + row;
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[32];
      tmem_ld32(tlane + 128u + (uint32_t)(ch * 32), v);
      if (valid) {
        float o[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] = __uint_as_float(v[c]) * inv;
        if (p.out) {
          float4* dst = reinterpret_cast<float4*>(p.out + grow * p.ldo + head * HD + ch * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) dst[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
        if (p.oh) {
          split_store16(p.oh + grow * p.ldh + head * HD + ch * 32, p.ol + grow * p.ldh + head * HD + ch * 32, o);
          split_store16(p.oh + grow * p.ldh + head * HD + ch * 32 + 16, p.ol + grow * p.ldh + head * HD + ch * 32 + 16, o + 16);
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ATT_TMEM) : "memory");
  }
}

// planes [rows, ld] (columns col0 .. col0 + C) -> transposed planes [C, ldt] (ldt >= rows)
__global__ void k_transpose_planes(const __half* __restrict__ xh, const __half* __restrict__ xl, int ld, int col0, int64_t rows,
                                   __half* __restrict__ th, __half* __restrict__ tl, int64_t ldt) {
  __shared__ __half sh[32][34], sl[32][34];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 256 threads: ty 0..7
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i;
    sh[i][tx] = r < rows ? xh[r * ld + col0 + c0 + tx] : __float2half(0.f);
    sl[i][tx] = r < rows ? xl[r * ld + col0 + c0 + tx] : __float2half(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + tx;
    if (r < ldt) {
      th[(int64_t)(c0 + i) * ldt + r] = sh[tx][i];
      tl[(int64_t)(c0 + i) * ldt + r] = sl[tx][i];
    }
  }
}

}  // namespace

static std::atomic<int> g_attn_tc{-1};
bool attention_tc_enabled() {
  int v = g_attn_tc.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("SSB_ATTN_TC");
    v = e ? (atoi(e) != 0 ? 1 : 0) : 1;  // default on (validated on B200: tests/test_gpu_tc.py, profiles/r02_*attention*)
    g_attn_tc.store(v, std::memory_order_relaxed);
  }
  return v != 0 && tc_available();
}
int set_attention_tc_enabled(int on) {
  g_attn_tc.store(on ? 1 : 0, std::memory_order_relaxed);
  return on ? 1 : 0;
}

int transpose_planes(Ctx& ctx, const __half* xh, const __half* xl, int ld, int col0, int64_t rows, int C, __half* th, __half* tl,
                     int64_t ldt) {
  if (ctx.dry || rows == 0) return 0;
  SSB_CHECK(C % 32 == 0 && ldt >= rows, "transpose_planes: bad shape");
  dim3 grid((unsigned)((ldt + 31) / 32), (unsigned)(C / 32));
  k_transpose_planes<<<grid, 256, 0, ctx.stream>>>(xh, xl, ld, col0, rows, th, tl, ldt);
  SSB_CUDA(cudaGetLastError());
  ++g_launches;
  return 0;
}

int attention_tc(Ctx& ctx, const AttnTCArgs& a) {
  if (ctx.dry || a.B == 0 || a.max_q == 0) return 0;
  SSB_CHECK(a.heads * HD <= 256 && a.ldvt % 8 == 0 && a.ldq % 8 == 0 && a.ldk % 8 == 0, "attention_tc: bad layout");
  SSB_CHECK(a.out || (a.oh && a.ol), "attention_tc: no output");
  {
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev].load(std::memory_order_acquire)) {
      SSB_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
      configured[dev].store(true, std::memory_order_release);
    }
  }
  CUtensorMap mq_h, mq_l, mk_h, mk_l, mv_h, mv_l;
  if (make_act_map(&mq_h, a.Qh, a.rows_q, a.ldq, AQ)) return -1;
  if (make_act_map(&mq_l, a.Ql, a.rows_q, a.ldq, AQ)) return -1;
  if (make_act_map(&mk_h, a.Kh, a.rows_k, a.ldk, AK)) return -1;
  if (make_act_map(&mk_l, a.Kl, a.rows_k, a.ldk, AK)) return -1;
  if (make_act_map(&mv_h, a.Vth, (int64_t)a.heads * HD, (int)a.ldvt, HD)) return -1;
  if (make_act_map(&mv_l, a.Vtl, (int64_t)a.heads * HD, (int)a.ldvt, HD)) return -1;
  AttnTCParams p;
  p.utt_q = a.utt_q; p.utt_k = a.utt_k; p.qcol0 = a.qcol0; p.kcol0 = a.kcol0; p.keymask = a.keymask;
  p.c2 = a.scale * 1.4426950408889634f;
  p.out = a.out; p.ldo = a.ldo; p.oh = a.oh; p.ol = a.ol; p.ldh = a.ldh;
  dim3 grid((a.max_q + AQ - 1) / AQ, a.heads, a.B);
  attention_tc_kernel<<<grid, 256, ATT_SMEM, ctx.stream>>>(mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, p);
  SSB_CUDA(cudaGetLastError());
  ++g_launches;
  return 0;
}

}  // namespace ssb
