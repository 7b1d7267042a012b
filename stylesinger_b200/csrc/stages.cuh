#pragma once
#include "model.cuh"

namespace ssb {

int fft_blocks(Ctx& c, const FFT& f, const SeqDev& s, float* x, const float* keep, bool tc = false);
int run_encoder(Ctx& c, const Model& m, const SeqDev& sp, const int32_t* tok_g, const int32_t* note_g,
                const int32_t* type_g, const float* ndur_g, float* srcmask, float* enc_out);
int run_fft_decoder(Ctx& c, const Model& m, const SeqDev& sf, const float* dec_in, float* xd, bool tc);
int run_duration_predictor(Ctx& c, const Model& m, const SeqDev& sp, const float* dur_inp, const float* srcmask,
                           float* logdur, int32_t* dur);
int run_style(Ctx& c, const Model& m, const SeqDev& sf, const SeqDev& sr, const float* dec0, const float* ref_g,
              const float* reff0_g, float* style, int32_t* codes, float* rq_in_out);
struct DenoiserBufs {
  float *x, *y, *zg, *skip, *sbuf, *head, *condall;
  float* condpre;             // tensor-core path with the hoisted conditioner: [rows, L*2C] fp32 pre-activation addends
  __half *yh, *yl, *zh, *zl;  // tensor-core path: fp16 hi/lo planes of y = x + step bias and of the gate output
  __half *ch, *cl;            // tensor-core path: fp16 hi/lo planes of the conditioner [rows,256]
  __half *skh, *skl, *sh, *sl;  // tensor-core heads: planes of the skip sum and of relu(skip_proj)
  __half *x80h, *x80l;          // mel net, tensor-core in_proj: planes of x_t padded to 128 columns
  bool tc_heads;
  bool skip_tiled;              // skip accumulator in the chunk-tiled layout (EpiTC::skip_tiled)
  int ld_head;
  bool tc;
};
bool denoiser_tc_ok(const Model& m, const Denoiser& d);
int alloc_denoiser(Ctx& c, const Denoiser& d, const SeqDev& s, bool tc, DenoiserBufs* b, bool hoist = false);
int prepare_cond(Ctx& c, const Denoiser& d, const SeqDev& s, const float* cond_g, DenoiserBufs& b);
// split (0 = off): first row tile of the second utterance group; the residual layers then run as two software-pipelined lanes
// on the interleaved gate / 1x1 kernel (conv_gemm_tc_dual)
int mel_denoiser_eval(Ctx& c, const Denoiser& d, const SeqDev& s, int t, const float* x80, DenoiserBufs& b, int split = 0);
int denoiser_stack(Ctx& c, const Denoiser& d, const SeqDev& s, int t, DenoiserBufs& b, int split = 0);
int lane_split_tile(const Seq* host_seq);  // utterance boundary closest to half of the row tiles (0: no split possible)
// both F0/UV samplers in lock step, their residual layers interleaved on the dual kernel (large batches)
bool f0_dual_ok(const Model& m, const SeqDev& s);
int run_f0_diffusion_dual(Ctx& c, const Model& m, const SeqDev& s, const float* cond0, const float* cond1, const float* lo,
                          const float* hi, const float* const gnoise[2], const float* const unoise[2], uint64_t seed,
                          float* const z[2], int32_t* const uv[2]);
// host_seq (optional): the host-side layout of `s`; enables the experimental utterance grouping (SSB_MEL_GROUP_FRAMES)
int run_mel_diffusion(Ctx& c, const Model& m, const SeqDev& s, const float* cond_g, const float* coarse_g,
                      const float* noise, uint64_t seed, float* mel_tight, const Seq* host_seq = nullptr);
int run_mel_diffusion_plms(Ctx& c, const Model& m, const SeqDev& s, const float* cond_g, const float* coarse_g,
                           const float* q_noise, uint64_t seed, int interval, float* mel_tight);
int run_f0_diffusion(Ctx& c, const Model& m, int which, const SeqDev& s, const float* cond_g, const float* lo,
                     const float* hi, const float* gnoise, const float* unoise, uint64_t seed, float* z, int32_t* uv,
                     const Seq* host_seq = nullptr);  // host_seq: see run_mel_diffusion (SSB_F0_GROUP_FRAMES)
int run_f0_diffusion_pair_persistent(Ctx& c, const Model& m, const SeqDev& s, const float* cond0, const float* cond1,
                                     const float* lo, const float* hi, const float* const gnoise[2],
                                     const float* const unoise[2], uint64_t seed, float* const z[2], int32_t* const uv[2]);
bool f0_pair_persistent_ok(const Model& m, const SeqDev& s);
int run_vocoder(Ctx& c, const Vocoder& v, const Seq& seq, const float* mel_tight, const float* f0_tight,
                const float* rand_ini, const float* src_noise, uint64_t seed, float* wav_tight);
int denoiser_eval_api(Ctx& c, const Model& m, int which, const SeqDev& s, const float* x_tight, const int32_t* uv_tight,
                      int t, const float* cond_tight, float* out_tight);
int run_acoustic(Ctx& c, const Model& m, const ssb_acoustic_inputs& in, const ssb_acoustic_outputs& out, bool durations_only,
                 int32_t* dur_out, float* logdur_out);

}  // namespace ssb
