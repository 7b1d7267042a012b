// Packed (device-resident) weights of the acoustic model and the vocoder, plus the stage drivers.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/stylesinger_b200.h"
#include "attention.cuh"
#include "common.cuh"
#include "conv_gemm.cuh"
#include "conv_gemm_tc.cuh"
#include "ops.cuh"

namespace ssb {

// One dense operator: weights [taps][Cin][Npad] + bias [N]
struct Conv {
  float* W = nullptr;
  float* bias = nullptr;
  int taps = 1, Cin = 0, N = 0, Npad = 0, dil = 1, center = 0;
};

enum PackMode { PACK_PLAIN = 0, PACK_GATE_SIG_TANH = 1 /* DiffNet: [sigmoid C | tanh C] */,
                PACK_GATE_TANH_SIG = 2 /* WN: [tanh C | sigmoid C] */ };

struct HostTensor {
  const float* data = nullptr;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// Owns every cudaMalloc'ed weight buffer.
struct DevicePool {
  std::vector<void*> ptrs;
  ~DevicePool();
  float* upload(const std::vector<float>& h);
  float* alloc(size_t n);
  void release(void* p);  // free one buffer of the pool early (schedule tables replaced by ssb_model_set_schedule)
};

struct TensorMap {
  std::map<std::string, HostTensor> t;
  std::string missing;
  const HostTensor* get(const std::string& name, std::initializer_list<int64_t> shape = {});
};

struct FFTLayer {
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  Conv qkv, out, ffn1, ffn2;
  ConvTC ffn1_tc, ffn2_tc;  // the FFN (92 % of the block's FLOPs) on the tcgen05 path, used for long sequences
  ConvTC qkv_tc, out_tc;    // self-attention in/out projections on the same path
};
struct FFT {
  std::vector<FFTLayer> layers;
  float *ln_g = nullptr, *ln_b = nullptr;
  float* pos_alpha = nullptr;  // device scalar, null for the encoder
  int kernel = 9;
};

struct DenoiserLayer {
  Conv dil;    // k3 dilated, gate-interleaved columns, N = 2C
  Conv outp;   // 1x1, N = 2C ([res | skip])
  Conv dproj;  // diffusion_projection C -> C (used only to build the step-bias table)
  ConvTC dil_tc, outp_tc;  // tensor-core packing of dil / outp (ok == false when not eligible)
  ConvTC cond_tc;          // conditioner_projection (256 -> 2C, gate-interleaved) as the 2nd K segment of dil_tc
  float* bias_gate_tc = nullptr;  // dil bias + conditioner bias (packed column order)
};
struct Denoiser {
  int C = 0, L = 0, in_dims = 0, out_dims = 0, cycle = 4;
  bool ddiff = false;
  Conv in_proj;                 // mel: 80 -> C (relu).  ddiff: unused (in_w / in_b below)
  float *in_w = nullptr, *in_b = nullptr, *uv_emb = nullptr;  // ddiff input: [C/2], [C/2], [2][C/2]
  Conv mlp0, mlp2;
  std::vector<DenoiserLayer> layers;
  Conv cond_all;                // 256 -> L*2C, gate-interleaved per layer
  ConvTC cond_all_tc;           // the same stacked projection for the tcgen05 kernel (hoisted out of the T loop; no bias)
  Conv skip_proj, out_proj;
  // schedule-dependent (set by ssb_model_set_schedule)
  int T = 0;
  float* dtab = nullptr;        // [T][L][C] step bias per layer
  float* gtab = nullptr;        // [T][8]
  float* mtab = nullptr;        // [T][8] (ddiff only)
  std::vector<float> gtab_h;
  // persistent-sampler extras (mel net): every GEMM of a diffusion step on the tcgen05 path
  ConvTC in_tc;    // input_projection, K padded 80 -> 128
  ConvTC skip_tc;  // skip_projection with the 1/sqrt(L) skip scale folded into the weights
  ConvTC out_tc;   // output_projection, N padded 80 -> 256 (4 N-tiles of 64: one cluster)
  float* out_bias_pad = nullptr;
  float* skip_bias_pad = nullptr;
};

struct AlignLayer {
  Conv q, kv, out, lin1, lin2;
  ConvTC q_tc, kv_tc, out_tc, lin1_tc, lin2_tc;  // tcgen05 packing of the same projections (long batches)
  float *n1_g, *n1_b, *n2_g, *n2_b;
};

struct Model {
  DevicePool pool;
  ssb_hparams hp;
  // tables
  float* pos_table = nullptr; int pos_rows = 0;
  float* tok_emb = nullptr; int n_tokens = 0;
  float *note_emb = nullptr, *type_emb = nullptr, *dur_w = nullptr, *dur_b = nullptr;
  float* pitch_emb = nullptr;
  float *spec_min = nullptr, *spec_max = nullptr;
  Conv spk_proj, emo_proj;
  FFT enc, dec;
  Conv dp_conv[4]; float* dp_ln_g[4]; float* dp_ln_b[4]; Conv dp_lin; int dp_layers = 2;
  // style adaptor
  Conv wn_in[4], wn_rs[4];
  struct CB { float *ln_g, *ln_b; Conv c1, c2; } cb[10];
  float *cb_last_g = nullptr, *cb_last_b = nullptr;
  Conv cb_post;
  float* codebooks = nullptr; float* cb_norm2 = nullptr;  // [depth][n_embed][256], [depth][n_embed]
  Conv l1;
  AlignLayer align[2];
  Denoiser f0net[2];
  Denoiser melnet;
  Conv mel_out, ln_proj;
  float log_eps = 0.f;
  // auxiliary stream + fork/join events: lets the two independent F0 samplers overlap (created in build_model).
  // Calls on one model are therefore serialised with respect to these events (one in-flight forward per model).
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool persistent = true;  // single-launch persistent sampler for small batches (ssb_model_set_persistent)
  bool cond_hoist = true;   // tcgen05 per-launch path: conditioner projection computed once per call ([rows, L*2C] fp32) and
                            // added in the GATE epilogue, instead of being contracted inside every layer GEMM of every step
  bool persistent_groups = false;  // large batches: groups of <= 48 row tiles, one persistent launch each (mel sampler)
  bool use_tc = true;  // tcgen05 path for the denoiser layer GEMMs (ssb_model_set_tensor_cores)
  bool fft_tc = true;  // tcgen05 path for the decoder FFT blocks' FFN on long batches (ssb_model_set_fft_tensor_cores)
};

struct VocStage {
  Conv up;          // transposed conv as 3-tap conv, N = u * Cout
  int u = 1, Cout = 0;
  float *nc_w = nullptr, *nc_b = nullptr; int nc_s = 1;  // noise conv
  float* nc_wt = nullptr;                                 // the same weights as [K, C] (tiled kernel)
  struct RB { Conv c1[3], c2[3]; ConvTC c1_tc[3], c2_tc[3]; } rb[4];
  ConvTC up_tc;     // tensor-core packing of the transposed conv
  bool res_tc = false;  // all ResBlock convs of this stage are tensor-core eligible (C % 64 == 0)
  bool paired = false;  // C == 32: ResBlock convs packed as 64-channel convs over PAIRS of time steps (see pack.cu)
};
struct Vocoder {
  DevicePool pool;
  ssb_vocoder_config cfg;
  Conv pre, post;
  std::vector<VocStage> stages;
  int nk = 3;
  float *lin_w = nullptr, *lin_b = nullptr;
  bool nsf = true;
  bool use_tc = true;
};

// ---- packing helpers (pack.cu) -------------------------------------------------------------------
int pack_conv(DevicePool& pool, const HostTensor* w, const HostTensor* b, int dil, PackMode mode, Conv* out,
              const HostTensor* g = nullptr /* weight-norm g: w is v */);
int pack_conv_tc(DevicePool& pool, const HostTensor* w, int dil, PackMode mode, const float* packed_bias, ConvTC* out);
int pack_linear(DevicePool& pool, const HostTensor* w, const HostTensor* b, Conv* out, int row0 = 0, int nrows = -1);
int pack_conv_transpose(DevicePool& pool, const HostTensor* v, const HostTensor* g, const HostTensor* b, int u, Conv* out);
int build_model(TensorMap& tm, const ssb_hparams& hp, Model* m);
int build_vocoder(TensorMap& tm, const ssb_vocoder_config& cfg, Vocoder* v);
int set_schedule(Model* m, int which, int T, const float* step_emb, const float* gtab, const float* mtab,
                 cudaStream_t stream);

// ---- op helpers (stages.cu) ------------------------------------------------------------------------
ConvGemm make_gemm(const Conv& c, const SeqDev& s, const float* A, int lda);

}  // namespace ssb
