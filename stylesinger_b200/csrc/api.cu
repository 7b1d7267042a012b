// extern "C" boundary (include/stylesinger_b200.h) + the whole-model driver.
#include <stdlib.h>
#include <string.h>

#include "stages.cuh"

struct ssb_model { ssb::Model m; };
struct ssb_vocoder { ssb::Vocoder v; };

namespace ssb {

#define RUN(x)                 \
  do {                         \
    int rc_ = (x);             \
    if (rc_ != 0) return rc_;  \
  } while (0)
#define WS_OK(c) SSB_CHECK((c).dry || !(c).failed, "workspace too small")

static int32_t* alloc_rows_i32(Ctx& c, const SeqDev& s, int C = 1) {
  int32_t* p = c.alloc<int32_t>((size_t)s.rows * C);
  if (!c.dry && p && !c.failed) cudaMemsetAsync(p, 0, (size_t)s.rows * C * sizeof(int32_t), c.stream);
  return p;
}

// project per-utterance vectors (spk_embed_proj / emo_embed_proj, stylesinger.py:130-132)
static int project_vec(Ctx& c, const Conv& w, const float* in_tight, int B, float* out_tight) {
  Seq s1;
  int32_t offs[2] = {0, B};
  s1.build(offs, 1);
  const size_t mk = c.mark();
  SeqDev sd;
  RUN(upload_layout(c, s1, 1, &sd));
  float* a = alloc_rows(c, sd, 256);
  float* o = alloc_rows(c, sd, 256);
  WS_OK(c);
  RUN(pack_rows(c, sd, in_tight, 256, a, 256, 256));
  ConvGemm g = make_gemm(w, sd, a, 256);
  g.e.out = o; g.e.ldo = 256;
  RUN(conv_gemm(c, g));
  RUN(unpack_rows(c, sd, o, 256, out_tight, 256, 256));
  c.release(mk);
  return 0;
}

// StyleSinger.forward(infer=True) (stylesinger.py:119-187); durations_only stops after add_dur.
int run_acoustic(Ctx& c, const Model& m, const ssb_acoustic_inputs& in, const ssb_acoustic_outputs& out,
                 bool durations_only, int32_t* dur_out, float* logdur_out) {
  const int H = 256, B = in.B;
  SSB_CHECK(B >= 1 && in.ph_offsets && in.ref_offsets, "acoustic: bad batch description");
  SSB_CHECK(durations_only || in.frame_offsets, "acoustic: frame_offsets required");
  SSB_CHECK(durations_only || in.mel2ph || in.dur, "acoustic: need mel2ph or dur");
  Seq qp, qr, qf;
  qp.build(in.ph_offsets, B);
  qr.build(in.ref_offsets, B);
  SSB_CHECK(qp.maxlen + 2 <= m.pos_rows && qr.maxlen + 2 <= m.pos_rows, "sequence longer than __pos_table");
  SeqDev sp, sr, sf;
  RUN(upload_layout(c, qp, 1, &sp));
  RUN(upload_layout(c, qr, 1, &sr));
  int32_t* tok = alloc_rows_i32(c, sp);
  int32_t* note = alloc_rows_i32(c, sp);
  int32_t* ntype = alloc_rows_i32(c, sp);
  float* ndur = alloc_rows(c, sp, 1);
  float* srcmask = alloc_rows(c, sp, 1);
  float* enc = alloc_rows(c, sp, H);
  float* spk = c.alloc<float>((size_t)B * H);
  float* emo = c.alloc<float>((size_t)B * H);
  WS_OK(c);
  RUN(pack_rows_i32(c, sp, in.txt_tokens, tok));
  RUN(pack_rows_i32(c, sp, in.note, note));
  RUN(pack_rows_i32(c, sp, in.note_type, ntype));
  RUN(pack_rows(c, sp, in.note_dur, 1, ndur, 1, 1));
  RUN(project_vec(c, m.spk_proj, in.spk_embed, B, spk));
  RUN(project_vec(c, m.emo_proj, in.emo_embed, B, emo));
  if (out.spk_proj && !c.dry) SSB_CUDA(cudaMemcpyAsync(out.spk_proj, spk, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, c.stream));
  if (out.emo_proj && !c.dry) SSB_CUDA(cudaMemcpyAsync(out.emo_proj, emo, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, c.stream));
  RUN(run_encoder(c, m, sp, tok, note, ntype, ndur, srcmask, enc));
  if (out.encoder_out) RUN(unpack_rows(c, sp, enc, H, out.encoder_out, H, H));

  if (durations_only) {
    float* dinp = alloc_rows(c, sp, H);
    float* logdur = alloc_rows(c, sp, 1);
    int32_t* dur = alloc_rows_i32(c, sp);
    WS_OK(c);
    CombineArgs a;  // dur_inp = (encoder_out + spk + emo) * src_nonpadding (stylesinger.py:134-138)
    a.m[0] = enc; a.ldm[0] = H; a.v[0] = spk; a.v[1] = emo; a.rowmask = srcmask; a.out = dinp; a.ldo = H; a.C = H;
    RUN(combine_rows(c, sp, a));
    RUN(run_duration_predictor(c, m, sp, dinp, srcmask, logdur, dur));
    if (dur_out) RUN(unpack_rows_i32(c, sp, dur, dur_out));
    if (logdur_out) RUN(unpack_rows(c, sp, logdur, 1, logdur_out, 1, 1));
    return 0;
  }

  qf.build(in.frame_offsets, B);
  SSB_CHECK(qf.maxlen + 2 <= m.pos_rows, "frame sequence longer than __pos_table");
  RUN(upload_layout(c, qf, 1, &sf));
  int32_t* mel2ph = alloc_rows_i32(c, sf);
  int32_t* midi = alloc_rows_i32(c, sf);
  float* tgt = alloc_rows(c, sf, 1);
  float* dec0 = alloc_rows(c, sf, H);
  float* ref = alloc_rows(c, sr, 80);
  float* reff0 = alloc_rows(c, sr, 1);
  float* style = alloc_rows(c, sf, H);
  int32_t* codes = alloc_rows_i32(c, sr, m.hp.rq_depth);
  WS_OK(c);
  if (in.mel2ph) {
    RUN(pack_rows_i32(c, sf, in.mel2ph, mel2ph));
  } else {
    int32_t* dur = alloc_rows_i32(c, sp);
    WS_OK(c);
    RUN(pack_rows_i32(c, sp, in.dur, dur));
    RUN(length_regulate(c, sf, sp, dur, mel2ph));
  }
  if (out.mel2ph) RUN(unpack_rows_i32(c, sf, mel2ph, out.mel2ph));
  RUN(expand_states(c, sf, sp, mel2ph, enc, H, dec0, H, H, note, midi, tgt));
  RUN(pack_rows(c, sr, in.ref_mels, 80, ref, 80, 80));
  RUN(pack_rows(c, sr, in.ref_f0, 1, reff0, 1, 1));
  RUN(run_style(c, m, sf, sr, dec0, ref, reff0, style, codes, nullptr));
  if (out.style) RUN(unpack_rows(c, sf, style, H, out.style, H, H));
  if (out.rq_codes) {
    // codes are [rows, depth] int32: unpack column by column through the i32 row copier
    for (int d = 0; d < m.hp.rq_depth; ++d) RUN(unpack_cols_i32(c, sr, codes, m.hp.rq_depth, d, out.rq_codes));
  }

  // ---- pitch (inpaint_pitch, stylesinger.py:216-247)
  float* pitch_pred = alloc_rows(c, sf, 2);
  float* f0_denorm = alloc_rows(c, sf, 1);
  int32_t* pitch = alloc_rows_i32(c, sf);
  float* f0_in = nullptr;
  float* uv_in = nullptr;
  WS_OK(c);
  {
    const size_t mk = c.mark();
    float* za = alloc_rows(c, sf, 1);
    float* zs = alloc_rows(c, sf, 1);
    int32_t* uva = alloc_rows_i32(c, sf);
    int32_t* uvs = alloc_rows_i32(c, sf);
    WS_OK(c);
    if (in.f0) {
      // teacher forcing: the reference skips the samplers when f0 is passed (add_gmdiff_pitch :251-254)
      f0_in = alloc_rows(c, sf, 1);
      uv_in = alloc_rows(c, sf, 1);
      WS_OK(c);
      RUN(pack_rows(c, sf, in.f0, 1, f0_in, 1, 1));
      if (in.uv) RUN(pack_rows(c, sf, in.uv, 1, uv_in, 1, 1));
    } else {
      float* lo = alloc_rows(c, sf, 1);
      float* hi = alloc_rows(c, sf, 1);
      float* cond = alloc_rows(c, sf, H);
      float* cond2 = alloc_rows(c, sf, H);
      WS_OK(c);
      RUN(midi_clip_band(c, sf, midi, lo, hi));
      {
        CombineArgs a;  // pitch_inp_domain_agnostic = decoder_inp * tgt_nonpadding (:156)
        a.m[0] = dec0; a.ldm[0] = H; a.rowmask = tgt; a.out = cond; a.ldo = H; a.C = H;
        RUN(combine_rows(c, sf, a));
      }
      {
        CombineArgs a;  // (decoder_inp + spk + emo + style) * tgt_nonpadding (:157-162)
        a.m[0] = dec0; a.ldm[0] = H; a.m[1] = style; a.ldm[1] = H; a.v[0] = spk; a.v[1] = emo;
        a.rowmask = tgt; a.out = cond2; a.ldo = H; a.C = H;
        RUN(combine_rows(c, sf, a));
      }
      // The two F0/UV samplers are independent (stylesinger.py:223-225): run the second one on the model's
      // auxiliary stream so that their latency-bound dependent chains overlap.  Disjoint workspace regions.
      if (f0_pair_persistent_ok(m, sf)) {
        float* zz[2] = {za, zs};
        int32_t* uu[2] = {uva, uvs};
        RUN(run_f0_diffusion_pair_persistent(c, m, sf, cond, cond2, lo, hi, in.f0_gauss_noise, in.f0_unif_noise, in.seed, zz, uu));
      } else if (f0_dual_ok(m, sf)) {
        // large batches: both nets in lock step, their gate / residual GEMMs interleaved on the dual kernel
        float* zz[2] = {za, zs};
        int32_t* uu[2] = {uva, uvs};
        RUN(run_f0_diffusion_dual(c, m, sf, cond, cond2, lo, hi, in.f0_gauss_noise, in.f0_unif_noise, in.seed, zz, uu));
      } else {
      // Two streams only for small batches (latency-bound chains).  From ~8k frames on every GEMM fills the GPU on its
      // own, and the CTA-pair (cluster) kernels used there must not run concurrently with each other from two streams:
      // that combination hung on B200 (gpurun diag, round 1; root cause open - see DESIGN.md "known issues").
      // SSB_F0_FORK_ALWAYS=1 (diagnosis only): fork at any size, i.e. the round-1 configuration that hung with CTA-pair kernels
      static const bool fork_always = getenv("SSB_F0_FORK_ALWAYS") != nullptr;
      const bool fork = !c.dry && m.aux_stream != nullptr && (sf.ntiles <= 64 || fork_always);
      if (fork) {
        SSB_CUDA(cudaEventRecord(m.ev_fork, c.stream));
        SSB_CUDA(cudaStreamWaitEvent(m.aux_stream, m.ev_fork, 0));
      }
      const size_t off0 = c.off;
      RUN(run_f0_diffusion(c, m, 0, sf, cond, lo, hi, in.f0_gauss_noise[0], in.f0_unif_noise[0], in.seed, za, uva, &qf));
      c.off = c.high;  // keep sampler 0's buffers alive: sampler 1 allocates above them
      {
        Ctx c2 = c;
        if (fork) c2.stream = m.aux_stream;
        RUN(run_f0_diffusion(c2, m, 1, sf, cond2, lo, hi, in.f0_gauss_noise[1], in.f0_unif_noise[1], in.seed, zs, uvs, &qf));
        if (c2.high > c.high) c.high = c2.high;
        c.failed = c.failed || c2.failed;
      }
      if (fork) {
        SSB_CUDA(cudaEventRecord(m.ev_join, m.aux_stream));
        SSB_CUDA(cudaStreamWaitEvent(c.stream, m.ev_join, 0));
      }
      c.off = off0;
      }
    }
    PitchGlueArgs pg;
    pg.za = za; pg.uva = uva; pg.zs = zs; pg.uvs = uvs; pg.midi = midi; pg.mel2ph = mel2ph;
    pg.f0_in = f0_in; pg.uv_in = in.uv ? uv_in : nullptr;
    pg.pitch_pred = pitch_pred; pg.f0_denorm = f0_denorm; pg.pitch = pitch;
    RUN(pitch_glue(c, sf, pg));
    c.release(mk);
  }
  if (out.pitch_pred) RUN(unpack_rows(c, sf, pitch_pred, 2, out.pitch_pred, 2, 2));
  if (out.f0_denorm) RUN(unpack_rows(c, sf, f0_denorm, 1, out.f0_denorm, 1, 1));

  // ---- decoder input (:165-172)
  float* dec = alloc_rows(c, sf, H);
  float* pemb = alloc_rows(c, sf, H);
  WS_OK(c);
  RUN(embed_rows(c, sf, pitch, m.pitch_emb, 300, 1.0f, pemb, H, H, 0));
  {
    CombineArgs a;
    a.m[0] = dec0; a.ldm[0] = H; a.m[1] = pemb; a.ldm[1] = H; a.m[2] = style; a.ldm[2] = H;
    a.v[0] = spk; a.v[1] = emo; a.rowmask = tgt; a.out = dec; a.ldo = H; a.C = H;
    RUN(combine_rows(c, sf, a));
  }
  if (out.decoder_inp) RUN(unpack_rows(c, sf, dec, H, out.decoder_inp, H, H));

  // ---- FFT decoder + mel_out (fs2.py:233-237, tts_modules.py:281-306)
  float* coarse = alloc_rows(c, sf, 80);
  float* cond = alloc_rows(c, sf, H);
  WS_OK(c);
  {
    const size_t mk = c.mark();
    float* xd = alloc_rows(c, sf, H);
    WS_OK(c);
    // long batches: the decoder's FFN GEMMs on the tcgen05 kernel (short ones stay on the fp32 FFMA path, which is
    // what the reference-golden parity tests pin)
    RUN(run_fft_decoder(c, m, sf, dec, xd, m.use_tc && m.fft_tc && tc_available() && sf.ntiles >= 8));
    {
      ConvGemm g = make_gemm(m.mel_out, sf, xd, H);
      g.e.rowmask = tgt; g.e.out = coarse; g.e.ldo = 80;
      RUN(conv_gemm(c, g));
    }
    // run_diffsinger: g = ln_proj(cat[coarse, decoder_inp, spk, emo, style]) (stylesinger.py:313-327)
    float* cat = alloc_rows(c, sf, 1104);
    WS_OK(c);
    RUN(concat_cond(c, sf, coarse, dec, spk, emo, style, cat));
    {
      ConvGemm g = make_gemm(m.ln_proj, sf, cat, 1104);
      g.e.out = cond; g.e.ldo = H;
      RUN(conv_gemm(c, g));
    }
    c.release(mk);
  }
  if (out.coarse_mel) RUN(unpack_rows(c, sf, coarse, 80, out.coarse_mel, 80, 80));
  if (out.diff_cond) RUN(unpack_rows(c, sf, cond, H, out.diff_cond, H, H));
  if (!in.skip_mel_diffusion) {
    SSB_CHECK(out.mel_out != nullptr, "acoustic: mel_out required");
    if (in.pndm_speedup > 0)  // PLMS: mel_noise, when given, supplies only the q_sample draw (its first [sumF, 80] block)
      RUN(run_mel_diffusion_plms(c, m, sf, cond, coarse, in.mel_noise, in.seed, in.pndm_speedup, out.mel_out));
    else
      RUN(run_mel_diffusion(c, m, sf, cond, coarse, in.mel_noise, in.seed, out.mel_out, &qf));
  }
  return 0;
}

int denoiser_eval_api(Ctx& c, const Model& m, int which, const SeqDev& s, const float* x_tight, const int32_t* uv_tight,
                      int t, const float* cond_tight, float* out_tight) {
  const Denoiser& d = which == 0 ? m.melnet : m.f0net[which - 1];
  SSB_CHECK(d.T > 0, "denoiser_eval: schedule not set");
  float* cond = alloc_rows(c, s, 256);
  DenoiserBufs b;
  RUN(alloc_denoiser(c, d, s, denoiser_tc_ok(m, d), &b, false));  // a single evaluation: nothing to amortise a hoist over
  RUN(pack_rows(c, s, cond_tight, 256, cond, 256, 256));
  RUN(prepare_cond(c, d, s, cond, b));
  if (which == 0) {
    float* x80 = alloc_rows(c, s, 80);
    WS_OK(c);
    RUN(pack_rows(c, s, x_tight, 80, x80, 80, 80));
    RUN(mel_denoiser_eval(c, d, s, t, x80, b));
  } else {
    float* z = alloc_rows(c, s, 1);
    int32_t* uv = alloc_rows_i32(c, s);
    WS_OK(c);
    RUN(pack_rows(c, s, x_tight, 1, z, 1, 1));
    RUN(pack_rows_i32(c, s, uv_tight, uv));
    const float* dt = d.dtab + (size_t)t * d.L * d.C;
    RUN(ddiff_input(c, s, z, uv, d.in_w, d.in_b, d.uv_emb, dt, b.tc ? nullptr : b.x, b.y, d.C, b.yh, b.yl));
    RUN(denoiser_stack(c, d, s, t, b));
  }
  RUN(unpack_rows(c, s, b.head, b.ld_head, out_tight, d.out_dims, d.out_dims));
  return 0;
}

}  // namespace ssb

// ================================================================================================
using namespace ssb;

static int to_map(const ssb_tensor_desc* t, int n, TensorMap* tm) {
  for (int i = 0; i < n; ++i) {
    SSB_CHECK(t[i].name && t[i].data && t[i].ndim >= 1 && t[i].ndim <= 4, "bad tensor descriptor");
    HostTensor h;
    h.data = t[i].data;
    for (int d = 0; d < t[i].ndim; ++d) h.shape.push_back(t[i].shape[d]);
    tm->t[t[i].name] = h;
  }
  return 0;
}
static Ctx make_ctx(void* ws, size_t bytes, void* stream, bool dry = false) {
  Ctx c;
  c.base = (char*)ws; c.cap = bytes; c.stream = (cudaStream_t)stream; c.dry = dry;
  return c;
}

extern "C" {

int ssb_version(void) { return 100; }
const char* ssb_last_error(void) { return ssb::last_error(); }

int ssb_model_create(ssb_model_t** out, const ssb_tensor_desc* tensors, int32_t n, const ssb_hparams* hp) {
  SSB_CHECK(out && tensors && hp, "ssb_model_create: null argument");
  TensorMap tm;
  if (to_map(tensors, n, &tm)) return -1;
  ssb_model* m = new ssb_model();
  if (build_model(tm, *hp, &m->m) != 0) {
    delete m;
    return -1;
  }
  *out = m;
  return 0;
}
void ssb_model_free(ssb_model_t* m) {
  if (!m) return;
  if (m->m.aux_stream) {
    cudaStreamSynchronize(m->m.aux_stream);
    cudaEventDestroy(m->m.ev_fork);
    cudaEventDestroy(m->m.ev_join);
    cudaStreamDestroy(m->m.aux_stream);
  }
  delete m;
}

int ssb_model_set_schedule(ssb_model_t* m, int32_t which, int32_t T, const float* step_emb, const float* gauss_tab,
                           const float* multi_tab, void* stream) {
  SSB_CHECK(m, "null model");
  return set_schedule(&m->m, which, T, step_emb, gauss_tab, multi_tab, (cudaStream_t)stream);
}

size_t ssb_durations_workspace_bytes(const ssb_model_t* m, const ssb_acoustic_inputs* in) {
  Ctx c = make_ctx(nullptr, 0, nullptr, true);
  ssb_acoustic_outputs o;
  memset(&o, 0, sizeof(o));
  if (run_acoustic(c, m->m, *in, o, true, nullptr, nullptr) != 0) return 0;
  return c.high + 4096;
}
int ssb_predict_durations(const ssb_model_t* m, const ssb_acoustic_inputs* in, int32_t* dur_out, float* logdur_out,
                          void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && in && workspace, "null argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  ssb_acoustic_outputs o;
  memset(&o, 0, sizeof(o));
  return run_acoustic(c, m->m, *in, o, true, dur_out, logdur_out);
}
size_t ssb_acoustic_workspace_bytes(const ssb_model_t* m, const ssb_acoustic_inputs* in) {
  Ctx c = make_ctx(nullptr, 0, nullptr, true);
  ssb_acoustic_outputs o;
  memset(&o, 0, sizeof(o));
  o.mel_out = (float*)(uintptr_t)256;
  if (run_acoustic(c, m->m, *in, o, false, nullptr, nullptr) != 0) return 0;
  return c.high + 4096;
}
int ssb_acoustic_forward(const ssb_model_t* m, const ssb_acoustic_inputs* in, const ssb_acoustic_outputs* out,
                         void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && in && out && workspace, "null argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  return run_acoustic(c, m->m, *in, *out, false, nullptr, nullptr);
}

static int mel_diff_impl(Ctx& c, const Model& m, const float* cond, const float* coarse, const int32_t* offs, int B,
                         const float* noise, uint64_t seed, float* mel_out) {
  Seq q;
  q.build(offs, B);
  SeqDev s;
  RUN(upload_layout(c, q, 1, &s));
  float* cg = alloc_rows(c, s, 256);
  float* co = alloc_rows(c, s, 80);
  WS_OK(c);
  RUN(pack_rows(c, s, cond, 256, cg, 256, 256));
  RUN(pack_rows(c, s, coarse, 80, co, 80, 80));
  return run_mel_diffusion(c, m, s, cg, co, noise, seed, mel_out, &q);
}
static int mel_plms_impl(Ctx& c, const Model& m, const float* cond, const float* coarse, const int32_t* offs, int B,
                         const float* q_noise, uint64_t seed, int interval, float* mel_out) {
  Seq q;
  q.build(offs, B);
  SeqDev s;
  RUN(upload_layout(c, q, 1, &s));
  float* cg = alloc_rows(c, s, 256);
  float* co = alloc_rows(c, s, 80);
  WS_OK(c);
  RUN(pack_rows(c, s, cond, 256, cg, 256, 256));
  RUN(pack_rows(c, s, coarse, 80, co, 80, 80));
  return run_mel_diffusion_plms(c, m, s, cg, co, q_noise, seed, interval, mel_out);
}
size_t ssb_mel_diffusion_plms_workspace_bytes(const ssb_model_t* m, const int32_t* frame_offsets, int32_t B) {
  Ctx c = make_ctx(nullptr, 0, nullptr, true);
  if (mel_plms_impl(c, m->m, nullptr, nullptr, frame_offsets, B, nullptr, 0, 1, nullptr) != 0) return 0;
  return c.high + 4096;
}
int ssb_mel_diffusion_sample_plms(const ssb_model_t* m, const float* cond, const float* coarse_mel,
                                  const int32_t* frame_offsets, int32_t B, const float* q_noise, uint64_t seed,
                                  int32_t interval, float* mel_out, void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && cond && coarse_mel && frame_offsets && mel_out && workspace, "null argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  return mel_plms_impl(c, m->m, cond, coarse_mel, frame_offsets, B, q_noise, seed, interval, mel_out);
}
size_t ssb_mel_diffusion_workspace_bytes(const ssb_model_t* m, const int32_t* frame_offsets, int32_t B) {
  Ctx c = make_ctx(nullptr, 0, nullptr, true);
  if (mel_diff_impl(c, m->m, nullptr, nullptr, frame_offsets, B, nullptr, 0, nullptr) != 0) return 0;
  return c.high + 4096;
}
int ssb_mel_diffusion_sample(const ssb_model_t* m, const float* cond, const float* coarse_mel,
                             const int32_t* frame_offsets, int32_t B, const float* noise, uint64_t seed,
                             float* mel_out, void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && cond && coarse_mel && frame_offsets && mel_out && workspace, "null argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  return mel_diff_impl(c, m->m, cond, coarse_mel, frame_offsets, B, noise, seed, mel_out);
}

int ssb_denoiser_eval(const ssb_model_t* m, int32_t which, const float* x, const int32_t* uv, int32_t t,
                      const float* cond, const int32_t* frame_offsets, int32_t B, float* out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && x && cond && frame_offsets && out && workspace && which >= 0 && which <= 2, "bad argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  Seq q;
  q.build(frame_offsets, B);
  SeqDev s;
  RUN(upload_layout(c, q, 1, &s));
  return denoiser_eval_api(c, m->m, which, s, x, uv, t, cond, out);
}

int ssb_f0_diffusion_sample(const ssb_model_t* m, int32_t which, const float* cond, const float* clip_lo,
                            const float* clip_hi, const int32_t* frame_offsets, int32_t B, const float* gauss_noise,
                            const float* unif_noise, uint64_t seed, float* f0_norm_out, int32_t* uv_out,
                            void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && cond && clip_lo && clip_hi && frame_offsets && f0_norm_out && uv_out && workspace, "null argument");
  SSB_CHECK(which == 0 || which == 1, "which must be 0 or 1");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  Seq q;
  q.build(frame_offsets, B);
  SeqDev s;
  RUN(upload_layout(c, q, 1, &s));
  float* cg = alloc_rows(c, s, 256);
  float* lo = alloc_rows(c, s, 1);
  float* hi = alloc_rows(c, s, 1);
  float* z = alloc_rows(c, s, 1);
  int32_t* uv = alloc_rows_i32(c, s);
  WS_OK(c);
  RUN(pack_rows(c, s, cond, 256, cg, 256, 256));
  RUN(pack_rows(c, s, clip_lo, 1, lo, 1, 1));
  RUN(pack_rows(c, s, clip_hi, 1, hi, 1, 1));
  RUN(run_f0_diffusion(c, m->m, which, s, cg, lo, hi, gauss_noise, unif_noise, seed, z, uv, &q));
  RUN(unpack_rows(c, s, z, 1, f0_norm_out, 1, 1));
  RUN(unpack_rows_i32(c, s, uv, uv_out));
  return 0;
}

// ---- per-registry drop-ins (FS_ENCODERS / FS_DECODERS 'fft', StyleSinger.get_style) ------------------------------------
static int fft_encoder_impl(Ctx& c, const Model& m, const int32_t* tokens, const int32_t* offs, int B, float* out) {
  Seq q;
  q.build(offs, B);
  SSB_CHECK(q.maxlen + 2 <= m.pos_rows, "sequence longer than __pos_table");
  SeqDev sp;
  RUN(upload_layout(c, q, 1, &sp));
  int32_t* tok = alloc_rows_i32(c, sp);
  float* srcmask = alloc_rows(c, sp, 1);
  float* enc = alloc_rows(c, sp, 256);
  WS_OK(c);
  RUN(pack_rows_i32(c, sp, tokens, tok));
  RUN(run_encoder(c, m, sp, tok, nullptr, nullptr, nullptr, srcmask, enc));
  return unpack_rows(c, sp, enc, 256, out, 256, 256);
}
static int fft_decoder_impl(Ctx& c, const Model& m, const float* x, const int32_t* offs, int B, float* out) {
  Seq q;
  q.build(offs, B);
  SSB_CHECK(q.maxlen + 2 <= m.pos_rows, "frame sequence longer than __pos_table");
  SeqDev sf;
  RUN(upload_layout(c, q, 1, &sf));
  float* xin = alloc_rows(c, sf, 256);
  float* xd = alloc_rows(c, sf, 256);
  WS_OK(c);
  RUN(pack_rows(c, sf, x, 256, xin, 256, 256));
  RUN(run_fft_decoder(c, m, sf, xin, xd, m.use_tc && m.fft_tc && tc_available() && sf.ntiles >= 8));
  return unpack_rows(c, sf, xd, 256, out, 256, 256);
}
static int get_style_impl(Ctx& c, const Model& m, const float* dec_inp, const int32_t* foffs, const float* ref_mels,
                          const float* ref_f0, const int32_t* roffs, int B, float* style_out, int32_t* codes_out) {
  Seq qf, qr;
  qf.build(foffs, B);
  qr.build(roffs, B);
  SSB_CHECK(qf.maxlen + 2 <= m.pos_rows && qr.maxlen + 2 <= m.pos_rows, "sequence longer than __pos_table");
  SeqDev sf, sr;
  RUN(upload_layout(c, qf, 1, &sf));
  RUN(upload_layout(c, qr, 1, &sr));
  float* dec0 = alloc_rows(c, sf, 256);
  float* style = alloc_rows(c, sf, 256);
  float* ref = alloc_rows(c, sr, 80);
  float* reff0 = alloc_rows(c, sr, 1);
  int32_t* codes = alloc_rows_i32(c, sr, m.hp.rq_depth);
  WS_OK(c);
  RUN(pack_rows(c, sf, dec_inp, 256, dec0, 256, 256));
  RUN(pack_rows(c, sr, ref_mels, 80, ref, 80, 80));
  RUN(pack_rows(c, sr, ref_f0, 1, reff0, 1, 1));
  RUN(run_style(c, m, sf, sr, dec0, ref, reff0, style, codes, nullptr));
  RUN(unpack_rows(c, sf, style, 256, style_out, 256, 256));
  if (codes_out)
    for (int d = 0; d < m.hp.rq_depth; ++d) RUN(unpack_cols_i32(c, sr, codes, m.hp.rq_depth, d, codes_out));
  return 0;
}

size_t ssb_fft_workspace_bytes(const ssb_model_t* m, int32_t which, const int32_t* offsets, int32_t B) {
  Ctx c = make_ctx(nullptr, 0, nullptr, true);
  const int rc = which == 0 ? fft_encoder_impl(c, m->m, nullptr, offsets, B, nullptr)
                            : fft_decoder_impl(c, m->m, nullptr, offsets, B, nullptr);
  return rc == 0 ? c.high + 4096 : 0;
}
int ssb_fft_encoder(const ssb_model_t* m, const int32_t* txt_tokens, const int32_t* ph_offsets, int32_t B, float* out,
                    void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && txt_tokens && ph_offsets && out && workspace, "null argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  return fft_encoder_impl(c, m->m, txt_tokens, ph_offsets, B, out);
}
int ssb_fft_decoder(const ssb_model_t* m, const float* x, const int32_t* frame_offsets, int32_t B, float* out,
                    void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && x && frame_offsets && out && workspace, "null argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  return fft_decoder_impl(c, m->m, x, frame_offsets, B, out);
}
size_t ssb_get_style_workspace_bytes(const ssb_model_t* m, const int32_t* frame_offsets, const int32_t* ref_offsets, int32_t B) {
  Ctx c = make_ctx(nullptr, 0, nullptr, true);
  return get_style_impl(c, m->m, nullptr, frame_offsets, nullptr, nullptr, ref_offsets, B, nullptr, nullptr) == 0 ? c.high + 4096 : 0;
}
int ssb_get_style(const ssb_model_t* m, const float* decoder_inp, const int32_t* frame_offsets, const float* ref_mels,
                  const float* ref_f0, const int32_t* ref_offsets, int32_t B, float* style_out, int32_t* codes_out,
                  void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && decoder_inp && frame_offsets && ref_mels && ref_f0 && ref_offsets && style_out && workspace, "null argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  return get_style_impl(c, m->m, decoder_inp, frame_offsets, ref_mels, ref_f0, ref_offsets, B, style_out, codes_out);
}

int ssb_rvq_lookup(const ssb_model_t* m, const float* x, const int32_t* ref_offsets, int32_t B, float* quant_out,
                   int32_t* codes_out, void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(m && x && ref_offsets && quant_out && codes_out && workspace, "null argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  Seq q;
  q.build(ref_offsets, B);
  SeqDev s;
  RUN(upload_layout(c, q, 1, &s));
  const int D = m->m.hp.rq_depth;
  float* xg = alloc_rows(c, s, 256);
  float* zq = alloc_rows(c, s, 256);
  int32_t* codes = alloc_rows_i32(c, s, D);
  WS_OK(c);
  RUN(pack_rows(c, s, x, 256, xg, 256, 256));
  RUN(rvq_lookup(c, s, xg, 256, m->m.codebooks, m->m.cb_norm2, m->m.hp.n_rq, D, zq, 256, codes));
  RUN(unpack_rows(c, s, zq, 256, quant_out, 256, 256));
  for (int d = 0; d < D; ++d) RUN(unpack_cols_i32(c, s, codes, D, d, codes_out));
  return 0;
}

int ssb_vocoder_create(ssb_vocoder_t** out, const ssb_tensor_desc* tensors, int32_t n, const ssb_vocoder_config* cfg) {
  SSB_CHECK(out && tensors && cfg, "ssb_vocoder_create: null argument");
  TensorMap tm;
  if (to_map(tensors, n, &tm)) return -1;
  ssb_vocoder* v = new ssb_vocoder();
  if (build_vocoder(tm, *cfg, &v->v) != 0) {
    delete v;
    return -1;
  }
  *out = v;
  return 0;
}
void ssb_vocoder_free(ssb_vocoder_t* v) { delete v; }

size_t ssb_vocoder_workspace_bytes(const ssb_vocoder_t* v, const int32_t* frame_offsets, int32_t B) {
  Ctx c = make_ctx(nullptr, 0, nullptr, true);
  Seq q;
  q.build(frame_offsets, B);
  if (run_vocoder(c, v->v, q, nullptr, (const float*)(uintptr_t)256, nullptr, nullptr, 0, nullptr) != 0) return 0;
  return c.high + 4096;
}
int ssb_hifigan_generate(const ssb_vocoder_t* v, const float* mel, const float* f0, const int32_t* frame_offsets,
                         int32_t B, const float* rand_ini, const float* src_noise, uint64_t seed, float* wav_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
  SSB_CHECK(v && mel && frame_offsets && wav_out && workspace, "null argument");
  Ctx c = make_ctx(workspace, workspace_bytes, stream);
  Seq q;
  q.build(frame_offsets, B);
  return run_vocoder(c, v->v, q, mel, f0, rand_ini, src_noise, seed, wav_out);
}

int ssb_model_set_tensor_cores(ssb_model_t* m, int32_t enable) {
  SSB_CHECK(m, "null model");
  m->m.use_tc = enable != 0 && tc_available();
  return m->m.use_tc ? 1 : 0;
}

int ssb_vocoder_set_tensor_cores(ssb_vocoder_t* v, int32_t enable) {
  SSB_CHECK(v, "null vocoder");
  v->v.use_tc = enable != 0 && tc_available();
  return v->v.use_tc ? 1 : 0;
}

int ssb_model_set_fft_tensor_cores(ssb_model_t* m, int32_t enable) {
  SSB_CHECK(m, "null model");
  m->m.fft_tc = enable != 0;
  return m->m.fft_tc ? 1 : 0;
}

int ssb_model_set_persistent(ssb_model_t* m, int32_t enable) {
  SSB_CHECK(m, "null model");
  m->m.persistent = enable != 0;
  return m->m.persistent ? 1 : 0;
}

int ssb_model_set_cond_hoist(ssb_model_t* m, int32_t enable) {
  SSB_CHECK(m, "null model");
  m->m.cond_hoist = enable != 0;
  return m->m.cond_hoist ? 1 : 0;
}

int ssb_model_set_persistent_groups(ssb_model_t* m, int32_t enable) {
  SSB_CHECK(m, "null model");
  m->m.persistent_groups = enable != 0;
  return m->m.persistent_groups ? 1 : 0;
}

int ssb_op_conv1d_tc(const float* x, const int32_t* offsets, int32_t B, int32_t Cin, const float* w_host,
                     const float* b_host, int32_t N, int32_t k, int32_t dilation, float* out, void* stream) {
  SSB_CHECK(x && offsets && w_host && out, "null argument");
  SSB_CHECK(tc_available(), "tensor-core path unavailable (cuTensorMapEncodeTiled)");
  DevicePool pool;
  HostTensor w, b;
  w.data = w_host; w.shape = {N, Cin, k};
  b.data = b_host; b.shape = {N};
  Conv cv;
  ConvTC ct;
  if (pack_conv(pool, &w, b_host ? &b : nullptr, dilation, PACK_PLAIN, &cv)) return -1;
  if (pack_conv_tc(pool, &w, dilation, PACK_PLAIN, cv.bias, &ct)) return -1;
  SSB_CHECK(ct.ok, "shape not eligible for the tensor-core path (Cin % 64, N % 128)");
  Seq q;
  q.build(offsets, B);
  const size_t bytes = ((size_t)q.rows() * (2 * Cin + N + 8) + 8 * (size_t)q.ntiles() + 1024) * sizeof(float) + (1 << 16);
  void* ws = nullptr;
  SSB_CUDA(cudaMalloc(&ws, bytes));
  Ctx c = make_ctx(ws, bytes, stream);
  SeqDev s;
  int rc = upload_layout(c, q, 1, &s);
  float* xg = alloc_rows(c, s, Cin);
  float* og = alloc_rows(c, s, N);
  __half* xh = c.alloc<__half>((size_t)s.rows * Cin);
  __half* xl = c.alloc<__half>((size_t)s.rows * Cin);
  if (rc == 0 && c.failed) rc = -1;
  if (rc == 0) rc = pack_rows(c, s, x, Cin, xg, Cin, Cin);
  if (rc == 0) rc = split_planes(c, xg, Cin, s.rows, Cin, 1.0f, xh, xl);
  if (rc == 0) {
    GemmTC g;
    g.A_hi = xh; g.A_lo = xl; g.rows_total = s.rows; g.w = &ct; g.tiles = s.tiles; g.ntiles = s.ntiles;
    g.e.mode = EPI_GENERIC; g.e.out = og; g.e.ldo = N;
    rc = conv_gemm_tc(c, g);
  }
  if (rc == 0) rc = unpack_rows(c, s, og, N, out, N, N);
  cudaError_t se = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(ws);
  if (rc == 0 && se != cudaSuccess) {
    ssb::set_error(std::string("ssb_op_conv1d_tc: ") + cudaGetErrorString(se));
    rc = -2;
  }
  return rc;
}

int ssb_mel_postprocess(float* mel, int64_t n_frames, float vmin, float vmax, int32_t* nonzero_frames, void* stream) {
  SSB_CHECK(mel && nonzero_frames && n_frames >= 0, "bad argument");
  return mel_postprocess_flat((cudaStream_t)stream, mel, n_frames, vmin, vmax, nonzero_frames);
}
int64_t ssb_launch_count(void) { return (int64_t)ssb::g_launches.load(); }
int32_t ssb_set_interleaved_layers(int32_t enable) { return ssb::set_dual_enabled(enable); }
int32_t ssb_set_attention_tensor_cores(int32_t enable) { return ssb::set_attention_tc_enabled(enable); }
int64_t ssb_variant_launch_count(const char* variant) { return variant ? (int64_t)ssb::variant_launch_count(variant) : 0; }
int32_t ssb_variant_names(char* buf, int32_t cap) { return buf && cap > 0 ? ssb::variant_names(buf, cap) : 0; }
void ssb_tensor_map_cache_stats(int64_t* encodes, int64_t* hits) {
  long long e = 0, h = 0;
  ssb::tensor_map_cache_stats(&e, &h);
  if (encodes) *encodes = e;
  if (hits) *hits = h;
}

int ssb_op_conv1d(const float* x, const int32_t* offsets, int32_t B, int32_t Cin, const float* w_host,
                  const float* b_host, int32_t N, int32_t k, int32_t dilation, int32_t act, float* out, void* stream) {
  SSB_CHECK(x && offsets && w_host && out, "null argument");
  DevicePool pool;
  HostTensor w, b;
  w.data = w_host; w.shape = {N, Cin, k};
  b.data = b_host; b.shape = {N};
  Conv cv;
  if (pack_conv(pool, &w, b_host ? &b : nullptr, dilation, PACK_PLAIN, &cv)) return -1;
  Seq q;
  q.build(offsets, B);
  const size_t bytes = ((size_t)q.rows() * (Cin + N + 8) + 8 * (size_t)q.ntiles() + 1024) * sizeof(float) + (1 << 16);
  void* ws = nullptr;
  SSB_CUDA(cudaMalloc(&ws, bytes));
  Ctx c = make_ctx(ws, bytes, stream);
  int rc = 0;
  SeqDev s;
  rc = upload_layout(c, q, 1, &s);
  float* xg = alloc_rows(c, s, Cin);
  float* og = alloc_rows(c, s, N);
  if (rc == 0 && c.failed) rc = -1;
  if (rc == 0) rc = pack_rows(c, s, x, Cin, xg, Cin, Cin);
  if (rc == 0) {
    ConvGemm g = make_gemm(cv, s, xg, Cin);
    g.e.act = act; g.e.out = og; g.e.ldo = N;
    rc = conv_gemm(c, g);
  }
  if (rc == 0) rc = unpack_rows(c, s, og, N, out, N, N);
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(ws);
  return rc;
}

int ssb_op_attention(const float* q, const float* k, const float* v, const int32_t* q_offsets,
                     const int32_t* k_offsets, int32_t B, float scale, float* out, void* stream) {
  SSB_CHECK(q && k && v && q_offsets && k_offsets && out, "null argument");
  Seq sq, sk;
  sq.build(q_offsets, B);
  sk.build(k_offsets, B);
  const size_t bytes = ((size_t)sq.rows() * 512 + (size_t)sk.rows() * 512 + 4096) * sizeof(float) + (1 << 16);
  void* ws = nullptr;
  SSB_CUDA(cudaMalloc(&ws, bytes));
  Ctx c = make_ctx(ws, bytes, stream);
  SeqDev dq, dk;
  int rc = upload_layout(c, sq, 1, &dq);
  if (rc == 0) rc = upload_layout(c, sk, 1, &dk);
  float* qg = alloc_rows(c, dq, 256);
  float* og = alloc_rows(c, dq, 256);
  float* kg = alloc_rows(c, dk, 256);
  float* vg = alloc_rows(c, dk, 256);
  if (rc == 0 && c.failed) rc = -1;
  if (rc == 0) rc = pack_rows(c, dq, q, 256, qg, 256, 256);
  if (rc == 0) rc = pack_rows(c, dk, k, 256, kg, 256, 256);
  if (rc == 0) rc = pack_rows(c, dk, v, 256, vg, 256, 256);
  if (rc == 0) {
    AttnArgs a;
    a.utt_q = dq.utt; a.utt_k = dk.utt; a.B = B; a.max_q = dq.maxlen; a.heads = 2;
    a.Q = qg; a.ldq = 256; a.K = kg; a.ldk = 256; a.V = vg; a.ldv = 256; a.scale = scale; a.out = og; a.ldo = 256;
    rc = attention(c, a);
  }
  if (rc == 0) rc = unpack_rows(c, dq, og, 256, out, 256, 256);
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(ws);
  return rc;
}

/* tcgen05 attention kernel at unit-test granularity: same contract as ssb_op_attention (fp32 in / out, tight rows) */
int ssb_op_attention_tc(const float* q, const float* k, const float* v, const int32_t* q_offsets,
                        const int32_t* k_offsets, int32_t B, float scale, float* out, void* stream) {
  SSB_CHECK(q && k && v && q_offsets && k_offsets && out, "null argument");
  SSB_CHECK(tc_available(), "tensor-core path unavailable (cuTensorMapEncodeTiled)");
  Seq sq, sk;
  sq.build(q_offsets, B);
  sk.build(k_offsets, B);
  const int64_t ldvt = (sk.rows() + 7) & ~int64_t(7);
  const size_t bytes = ((size_t)sq.rows() * 768 + (size_t)sk.rows() * 1024 + (size_t)ldvt * 256 + 4096) * sizeof(float) + (1 << 16);
  void* ws = nullptr;
  SSB_CUDA(cudaMalloc(&ws, bytes));
  Ctx c = make_ctx(ws, bytes, stream);
  SeqDev dq, dk;
  int rc = upload_layout(c, sq, 1, &dq);
  if (rc == 0) rc = upload_layout(c, sk, 1, &dk);
  float* qg = alloc_rows(c, dq, 256);
  float* og = alloc_rows(c, dq, 256);
  float* kg = alloc_rows(c, dk, 256);
  float* vg = alloc_rows(c, dk, 256);
  __half* pl[6];  // q, k, v planes (hi, lo)
  for (int i = 0; i < 6; ++i) pl[i] = c.alloc<__half>((size_t)(i < 2 ? dq.rows : dk.rows) * 256);
  __half* vth = c.alloc<__half>((size_t)ldvt * 256);
  __half* vtl = c.alloc<__half>((size_t)ldvt * 256);
  if (rc == 0 && c.failed) rc = -1;
  if (rc == 0) rc = pack_rows(c, dq, q, 256, qg, 256, 256);
  if (rc == 0) rc = pack_rows(c, dk, k, 256, kg, 256, 256);
  if (rc == 0) rc = pack_rows(c, dk, v, 256, vg, 256, 256);
  if (rc == 0) rc = split_planes(c, qg, 256, dq.rows, 256, 1.0f, pl[0], pl[1]);
  if (rc == 0) rc = split_planes(c, kg, 256, dk.rows, 256, 1.0f, pl[2], pl[3]);
  if (rc == 0) rc = split_planes(c, vg, 256, dk.rows, 256, 1.0f, pl[4], pl[5]);
  if (rc == 0) rc = transpose_planes(c, pl[4], pl[5], 256, 0, dk.rows, 256, vth, vtl, ldvt);
  if (rc == 0) {
    AttnTCArgs a;
    a.utt_q = dq.utt; a.utt_k = dk.utt; a.B = B; a.max_q = dq.maxlen; a.heads = 2;
    a.Qh = pl[0]; a.Ql = pl[1]; a.rows_q = dq.rows; a.ldq = 256;
    a.Kh = pl[2]; a.Kl = pl[3]; a.rows_k = dk.rows; a.ldk = 256;
    a.Vth = vth; a.Vtl = vtl; a.ldvt = ldvt;
    a.scale = scale; a.out = og; a.ldo = 256;
    rc = attention_tc(c, a);
  }
  if (rc == 0) rc = unpack_rows(c, dq, og, 256, out, 256, 256);
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(ws);
  return rc;
}

}  // extern "C"
