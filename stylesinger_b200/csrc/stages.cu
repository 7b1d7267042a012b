// Stage drivers: compose the kernels into the reference's modules (one function per SURVEY.md §8a group).
#include <string.h>

#include "sampler_tc.cuh"
#include "stages.cuh"

namespace ssb {

ConvGemm make_gemm(const Conv& c, const SeqDev& s, const float* A, int lda) {
  ConvGemm g;
  g.A = A; g.lda = lda; g.Cin = c.Cin; g.taps = c.taps; g.dil = c.dil; g.center = c.center;
  g.W = c.W; g.N = c.N; g.Npad = c.Npad; g.tiles = s.tiles; g.ntiles = s.ntiles;
  g.e.bias = c.bias;
  return g;
}

int upload_layout(Ctx& c, const Seq& s, int rate, SeqDev* out) {
  const int nt = s.ntiles(rate);
  int2* tiles = c.alloc<int2>((size_t)nt + 1);
  int4* utt = c.alloc<int4>((size_t)s.B + 1);
  int* ttight = c.alloc<int>((size_t)nt + 1);
  out->tile_tight = ttight;
  out->tiles = tiles; out->ntiles = nt; out->utt = utt; out->B = s.B; out->rows = s.rows(rate);
  out->total = s.total * rate; out->maxlen = s.maxlen * rate; out->rate = rate;
  if (c.dry) return 0;
  SSB_CHECK(!c.failed && tiles && utt, "workspace too small (layout tables)");
  std::vector<int2> ht((size_t)nt + 1);
  std::vector<int4> hu((size_t)s.B + 1);
  std::vector<int> htt((size_t)nt + 1);
  int k = 0;
  int64_t tight = 0;
  for (int b = 0; b < s.B; ++b) {
    const int len = s.len[b] * rate;
    const int rs = s.rs[b] * rate;
    hu[b] = make_int4(rs, len, (int)tight, 0);
    tight += len;
    for (int t0 = 0; t0 < len; t0 += TILE_M) {
      htt[k] = (int)(tight - len) + t0;
      ht[k++] = make_int2(rs + t0, (len - t0) < TILE_M ? (len - t0) : TILE_M);
    }
  }
  SSB_CUDA(cudaMemcpyAsync(ttight, htt.data(), sizeof(int) * nt, cudaMemcpyHostToDevice, c.stream));
  SSB_CUDA(cudaMemcpyAsync(tiles, ht.data(), sizeof(int2) * nt, cudaMemcpyHostToDevice, c.stream));
  SSB_CUDA(cudaMemcpyAsync(utt, hu.data(), sizeof(int4) * s.B, cudaMemcpyHostToDevice, c.stream));
  // pageable-source async copies are staged before returning, so the host vectors may die here
  return 0;
}

float* alloc_rows(Ctx& c, const SeqDev& s, int C, bool zero) {
  float* p = c.alloc<float>((size_t)s.rows * C);
  if (!c.dry && p && !c.failed && zero) cudaMemsetAsync(p, 0, (size_t)s.rows * C * sizeof(float), c.stream);
  return p;
}
static int32_t* alloc_rows_i32(Ctx& c, const SeqDev& s, int C = 1) {
  int32_t* p = c.alloc<int32_t>((size_t)s.rows * C);
  if (!c.dry && p && !c.failed) cudaMemsetAsync(p, 0, (size_t)s.rows * C * sizeof(int32_t), c.stream);
  return p;
}

static __half* alloc_half_rows(Ctx& c, const SeqDev& s, int C);

#define RUN(x)                 \
  do {                         \
    int rc_ = (x);             \
    if (rc_ != 0) return rc_;  \
  } while (0)
#define WS_OK(c) SSB_CHECK((c).dry || !(c).failed, "workspace too small")

// ------------------------------------------------------------------------------------------------
// a2-a5: FFTBlocks body (tts_modules.py:293-305 + EncSALayer common_layers.py:649-673)
// x [rows,256] in/out; keep = 1 - padding_mask (row mask, also the key mask)
int fft_blocks(Ctx& c, const FFT& f, const SeqDev& s, float* x, const float* keep, bool tc) {
  const int H = 256;
  const size_t mk = c.mark();
  for (auto& L : f.layers) tc = tc && L.ffn1_tc.ok && L.ffn2_tc.ok && L.qkv_tc.ok && L.out_tc.ok;
  float* h = alloc_rows(c, s, H);
  float* qkv = alloc_rows(c, s, 3 * H);
  float* att = alloc_rows(c, s, H);
  float* ff = tc ? nullptr : alloc_rows(c, s, 4 * H);
  __half *hh = nullptr, *hl = nullptr, *fh = nullptr, *fl = nullptr;  // fp16 hi/lo planes of h and of gelu(ffn_1)
  // tcgen05 attention (attention_tc.cu): q/k/v only as fp16 hi/lo planes [rows, 768] + V^T planes [256, ldvt]
  const bool atc = tc && attention_tc_enabled();
  const int64_t ldvt = (s.rows + 7) & ~int64_t(7);
  __half *qh = nullptr, *ql = nullptr, *vth = nullptr, *vtl = nullptr;
  if (tc) {
    hh = alloc_half_rows(c, s, H); hl = alloc_half_rows(c, s, H);
    fh = alloc_half_rows(c, s, 4 * H); fl = alloc_half_rows(c, s, 4 * H);
  }
  if (atc) {
    qh = alloc_half_rows(c, s, 3 * H); ql = alloc_half_rows(c, s, 3 * H);
    vth = c.alloc<__half>((size_t)ldvt * H); vtl = c.alloc<__half>((size_t)ldvt * H);
  }
  WS_OK(c);
  for (size_t i = 0; i < f.layers.size(); ++i) {
    const FFTLayer& L = f.layers[i];
    RUN(layernorm_rows(c, s, x, H, h, H, H, L.ln1_g, L.ln1_b, 1e-5f, nullptr));
    if (tc) {
      RUN(split_planes(c, h, H, s.rows, H, 1.0f, hh, hl));
      GemmTC g;
      g.A_hi = hh; g.A_lo = hl; g.rows_total = s.rows; g.w = &L.qkv_tc; g.tiles = s.tiles; g.ntiles = s.ntiles;
      g.e.mode = EPI_GENERIC;
      if (atc) { g.e.oh = qh; g.e.ol = ql; g.e.ldh = 3 * H; }
      else { g.e.out = qkv; g.e.ldo = 3 * H; }
      RUN(conv_gemm_tc(c, g));
    } else {
      ConvGemm g = make_gemm(L.qkv, s, h, H);
      g.e.out = qkv; g.e.ldo = 3 * H;
      RUN(conv_gemm(c, g));
    }
    if (atc) {
      RUN(transpose_planes(c, qh, ql, 3 * H, 2 * H, s.rows, H, vth, vtl, ldvt));
      AttnTCArgs a;
      a.utt_q = s.utt; a.utt_k = s.utt; a.B = s.B; a.max_q = s.maxlen; a.heads = 2;
      a.Qh = qh; a.Ql = ql; a.rows_q = s.rows; a.ldq = 3 * H; a.qcol0 = 0;
      a.Kh = qh; a.Kl = ql; a.rows_k = s.rows; a.ldk = 3 * H; a.kcol0 = H;
      a.Vth = vth; a.Vtl = vtl; a.ldvt = ldvt;
      a.keymask = keep; a.scale = 0.08838834764831845f;  // 128^-0.5
      a.oh = hh; a.ol = hl; a.ldh = H;                   // straight into the out-projection's A planes
      RUN(attention_tc(c, a));
    } else {
      AttnArgs a;
      a.utt_q = s.utt; a.utt_k = s.utt; a.B = s.B; a.max_q = s.maxlen; a.heads = 2;
      a.Q = qkv; a.ldq = 3 * H; a.K = qkv + H; a.ldk = 3 * H; a.V = qkv + 2 * H; a.ldv = 3 * H;
      a.keymask = keep; a.scale = 0.08838834764831845f;  // 128^-0.5
      a.out = att; a.ldo = H;
      RUN(attention(c, a));
    }
    if (tc) {
      if (!atc) RUN(split_planes(c, att, H, s.rows, H, 1.0f, hh, hl));
      GemmTC g;
      g.A_hi = hh; g.A_lo = hl; g.rows_total = s.rows; g.w = &L.out_tc; g.tiles = s.tiles; g.ntiles = s.ntiles;
      g.e.mode = EPI_GENERIC; g.e.res = x; g.e.ld_res = H; g.e.rowmask = keep; g.e.out = x; g.e.ldo = H;
      RUN(conv_gemm_tc(c, g));
    } else {
      ConvGemm g = make_gemm(L.out, s, att, H);
      g.e.res = x; g.e.ld_res = H; g.e.rowmask = keep; g.e.out = x; g.e.ldo = H;
      RUN(conv_gemm(c, g));
    }
    RUN(layernorm_rows(c, s, x, H, h, H, H, L.ln2_g, L.ln2_b, 1e-5f, nullptr));
    if (tc) {  // TransformerFFNLayer (transformer.py): ffn_2(gelu(ffn_1(x) * k^-0.5)) on the tcgen05 kernel
      RUN(split_planes(c, h, H, s.rows, H, 1.0f, hh, hl));
      GemmTC g1;
      g1.A_hi = hh; g1.A_lo = hl; g1.rows_total = s.rows; g1.w = &L.ffn1_tc; g1.tiles = s.tiles; g1.ntiles = s.ntiles;
      g1.e.mode = EPI_GENERIC; g1.e.alpha = 1.0f / sqrtf((float)f.kernel); g1.e.act = ACT_GELU;
      g1.e.oh = fh; g1.e.ol = fl; g1.e.ldh = 4 * H;
      RUN(conv_gemm_tc(c, g1));
      GemmTC g2;
      g2.A_hi = fh; g2.A_lo = fl; g2.rows_total = s.rows; g2.w = &L.ffn2_tc; g2.tiles = s.tiles; g2.ntiles = s.ntiles;
      g2.e.mode = EPI_GENERIC; g2.e.res = x; g2.e.ld_res = H; g2.e.rowmask = keep; g2.e.out = x; g2.e.ldo = H;
      RUN(conv_gemm_tc(c, g2));
      continue;
    }
    {
      ConvGemm g = make_gemm(L.ffn1, s, h, H);
      g.e.alpha = 1.0f / sqrtf((float)f.kernel); g.e.act = ACT_GELU; g.e.out = ff; g.e.ldo = 4 * H;
      RUN(conv_gemm(c, g));
    }
    {
      ConvGemm g = make_gemm(L.ffn2, s, ff, 4 * H);
      g.e.res = x; g.e.ld_res = H; g.e.rowmask = keep; g.e.out = x; g.e.ldo = H;
      RUN(conv_gemm(c, g));
    }
  }
  // final LN * mask, in place via h
  RUN(layernorm_rows(c, s, x, H, h, H, H, f.ln_g, f.ln_b, 1e-5f, keep));
  if (!c.dry) SSB_CUDA(cudaMemcpyAsync(x, h, (size_t)s.rows * H * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  c.release(mk);
  return 0;
}

// a1 + a7: encoder_out = FastspeechEncoder(txt) + NoteEncoder(note, dur, type)
int run_encoder(Ctx& c, const Model& m, const SeqDev& sp, const int32_t* tok_g, const int32_t* note_g,
                const int32_t* type_g, const float* ndur_g, float* srcmask, float* enc_out) {
  const int H = 256;
  const size_t mk = c.mark();
  int32_t* pos = alloc_rows_i32(c, sp);
  WS_OK(c);
  RUN(token_nonzero_mask(c, sp, tok_g, srcmask));
  RUN(positions_from_mask(c, sp, srcmask, pos));
  RUN(embed_rows(c, sp, tok_g, m.tok_emb, m.n_tokens, 16.0f, enc_out, H, H, 0));
  RUN(add_positional(c, sp, enc_out, H, H, pos, m.pos_table, m.pos_rows, nullptr));
  // FFTBlocks.forward: x = x.transpose * nonpadding (tts_modules.py:293)
  {
    CombineArgs a;
    a.m[0] = enc_out; a.ldm[0] = H; a.rowmask = srcmask; a.out = enc_out; a.ldo = H; a.C = H;
    RUN(combine_rows(c, sp, a));
  }
  RUN(fft_blocks(c, m.enc, sp, enc_out, srcmask));
  // StyleSinger.forward: encoder_out + note_encoder(note, note_dur, note_type) (stylesinger.py:124-126); a null note_g
  // stops at FastspeechEncoder.forward (the FS_ENCODERS['fft'] drop-in, ssb_fft_encoder)
  if (note_g) RUN(note_encoder(c, sp, note_g, type_g, ndur_g, m.note_emb, m.type_emb, m.dur_w, m.dur_b, 16.0f, enc_out, H, H, 1));
  c.release(mk);
  return 0;
}

// a16 body: FastspeechDecoder.forward = FFTBlocks.forward(x) with the padding mask taken from x itself
// (tts_modules.py:281-306): xd [rows,256] in/out
int run_fft_decoder(Ctx& c, const Model& m, const SeqDev& sf, const float* dec_in, float* xd, bool tc) {
  const int H = 256;
  const size_t mk = c.mark();
  float* keep = alloc_rows(c, sf, 1);
  float* c0 = alloc_rows(c, sf, 1);
  int32_t* pos = alloc_rows_i32(c, sf);
  WS_OK(c);
  if (!c.dry && xd != dec_in)
    SSB_CUDA(cudaMemcpyAsync(xd, dec_in, (size_t)sf.rows * H * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  RUN(row_nonzero_mask(c, sf, dec_in, H, H, keep));
  RUN(col0_nonzero_mask(c, sf, dec_in, H, c0));
  RUN(positions_from_mask(c, sf, c0, pos));
  RUN(add_positional(c, sf, xd, H, H, pos, m.pos_table, m.pos_rows, m.dec.pos_alpha));
  {
    CombineArgs a;
    a.m[0] = xd; a.ldm[0] = H; a.rowmask = keep; a.out = xd; a.ldo = H; a.C = H;
    RUN(combine_rows(c, sf, a));
  }
  RUN(fft_blocks(c, m.dec, sf, xd, keep, tc));
  c.release(mk);
  return 0;
}

// a8: DurationPredictor.inference (tts_modules.py:105-130)
int run_duration_predictor(Ctx& c, const Model& m, const SeqDev& sp, const float* dur_inp, const float* srcmask,
                           float* logdur /*[rows]*/, int32_t* dur /*[rows]*/) {
  const int H = 256;
  const size_t mk = c.mark();
  float* a = alloc_rows(c, sp, H);
  float* b = alloc_rows(c, sp, H);
  WS_OK(c);
  const float* cur = dur_inp;
  for (int i = 0; i < m.dp_layers; ++i) {
    ConvGemm g = make_gemm(m.dp_conv[i], sp, cur, H);
    g.e.act = ACT_RELU; g.e.out = a; g.e.ldo = H;
    RUN(conv_gemm(c, g));
    RUN(layernorm_rows(c, sp, a, H, b, H, H, m.dp_ln_g[i], m.dp_ln_b[i], 1e-5f, srcmask));
    cur = b;  // next conv reads b and writes a; the LN after it overwrites b only once that conv has finished
  }
  {
    ConvGemm g = make_gemm(m.dp_lin, sp, cur, H);
    g.e.rowmask = srcmask; g.e.out = logdur; g.e.ldo = 1;
    RUN(conv_gemm(c, g));
  }
  RUN(dur_from_logits(c, sp, logdur, srcmask, dur));
  c.release(mk);
  return 0;
}

// a10-a12: get_style (stylesinger.py:189-214)
int run_style(Ctx& c, const Model& m, const SeqDev& sf, const SeqDev& sr, const float* dec0, const float* ref_g /*[rows,80]*/,
              const float* reff0_g /*[rows]*/, float* style /*[rows_f,256]*/, int32_t* codes /*[rows_r,depth] guarded*/,
              float* rq_in_out /*optional guarded [rows_r,256]*/) {
  const int H = 256;
  const size_t mk = c.mark();
  float* rmask = alloc_rows(c, sr, 1);
  float* x = alloc_rows(c, sr, 80);
  float* z = alloc_rows(c, sr, 80);
  float* wout = alloc_rows(c, sr, 80);
  float* h80 = alloc_rows(c, sr, 80);
  float* g160 = alloc_rows(c, sr, 160);
  float* np0 = alloc_rows(c, sr, 1);
  float* npi = alloc_rows(c, sr, 1);
  float* st = alloc_rows(c, sr, H);
  float* zq = alloc_rows(c, sr, H);
  float* cat = alloc_rows(c, sr, 2 * H);
  float* zl = alloc_rows(c, sr, H);
  float* kmask = alloc_rows(c, sr, 1);
  int32_t* pos = alloc_rows_i32(c, sr);
  float* kv = alloc_rows(c, sr, 2 * H);
  float* q = alloc_rows(c, sf, H);
  float* att = alloc_rows(c, sf, H);
  float* tmp = alloc_rows(c, sf, H);
  // long batches: the aligner's five projections per layer on the tcgen05 kernel (the 256 -> 2048 -> 256 feed-forward is
  // 90 % of its FLOPs; its hidden activation then only exists as fp16 hi/lo planes); attention itself stays fp32
  bool tc = m.use_tc && m.fft_tc && tc_available() && sf.ntiles >= 8;
  for (int i = 0; i < 2; ++i)
    tc = tc && m.align[i].q_tc.ok && m.align[i].kv_tc.ok && m.align[i].out_tc.ok && m.align[i].lin1_tc.ok && m.align[i].lin2_tc.ok;
  float* hid = tc ? nullptr : alloc_rows(c, sf, 2048);
  WS_OK(c);
  // LocalStyleAdaptor.forward (lse.py:103-129)
  RUN(col0_nonzero_mask(c, sr, ref_g, 80, rmask));
  if (!c.dry) SSB_CUDA(cudaMemcpyAsync(x, ref_g, (size_t)sr.rows * 80 * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  for (int i = 0; i < 4; ++i) {  // WN.forward (wavenet.py:54-78), g=None
    {
      ConvGemm g = make_gemm(m.wn_in[i], sr, x, 80);
      g.e.mode = EPI_GATE; g.e.out = z; g.e.ldo = 80;
      RUN(conv_gemm(c, g));
    }
    ConvGemm g = make_gemm(m.wn_rs[i], sr, z, 80);
    if (i < 3) {
      g.e.mode = EPI_RES_SKIP; g.e.C = 80; g.e.res = x; g.e.ld_res = 80; g.e.beta = 1.0f; g.e.rowmask = rmask;
      g.e.out = x; g.e.ldo = 80; g.e.skip = wout; g.e.ld_skip = 80; g.e.skip_init = (i == 0);
    } else {
      g.e.out = wout; g.e.ldo = 80; g.e.accum = 1; g.e.gamma = 1.0f;
    }
    RUN(conv_gemm(c, g));
  }
  // ref_ph = wn_out * mask + ref_f0 broadcast (lse.py:110,121-123)
  RUN(scale_mask_add_rowscalar(c, sr, wout, 80, 80, rmask, reff0_g, x, 80));
  // ConvBlocks (lse.py:229-240)
  RUN(row_nonzero_mask(c, sr, x, 80, 80, np0));
  for (int i = 0; i < 5; ++i) {
    RUN(row_nonzero_mask(c, sr, x, 80, 80, npi));
    for (int j = 0; j < 2; ++j) {
      const Model::CB& b = m.cb[i * 2 + j];
      RUN(layernorm_rows(c, sr, x, 80, h80, 80, 80, b.ln_g, b.ln_b, 1e-5f, nullptr));
      {
        ConvGemm g = make_gemm(b.c1, sr, h80, 80);
        g.e.alpha = 1.0f / sqrtf(5.0f); g.e.act = ACT_GELU; g.e.out = g160; g.e.ldo = 160;
        RUN(conv_gemm(c, g));
      }
      {
        ConvGemm g = make_gemm(b.c2, sr, g160, 160);
        g.e.res = x; g.e.ld_res = 80; g.e.rowmask = npi; g.e.out = x; g.e.ldo = 80;
        RUN(conv_gemm(c, g));
      }
    }
  }
  {
    CombineArgs a;
    a.m[0] = x; a.ldm[0] = 80; a.rowmask = np0; a.out = x; a.ldo = 80; a.C = 80;
    RUN(combine_rows(c, sr, a));
  }
  RUN(layernorm_rows(c, sr, x, 80, h80, 80, 80, m.cb_last_g, m.cb_last_b, 1e-5f, np0));
  {
    ConvGemm g = make_gemm(m.cb_post, sr, h80, 80);
    g.e.rowmask = np0; g.e.out = st; g.e.ldo = H;
    RUN(conv_gemm(c, g));
  }
  if (rq_in_out && !c.dry)
    SSB_CUDA(cudaMemcpyAsync(rq_in_out, st, (size_t)sr.rows * H * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  RUN(rvq_lookup(c, sr, st, H, m.codebooks, m.cb_norm2, m.hp.n_rq, m.hp.rq_depth, zq, H, codes));
  // positions + l1 (stylesinger.py:198-200)
  RUN(col0_nonzero_mask(c, sr, zq, H, kmask));
  RUN(positions_from_mask(c, sr, kmask, pos));
  RUN(concat2_pos(c, sr, zq, H, pos, m.pos_table, m.pos_rows, cat, 2 * H));
  {
    ConvGemm g = make_gemm(m.l1, sr, cat, 2 * H);
    g.e.out = zl; g.e.ldo = H;
    RUN(conv_gemm(c, g));
  }
  RUN(col0_nonzero_mask(c, sr, zl, H, kmask));  // style_key_padding_mask = zl[:,:,0].eq(0) (:204) -> attend where != 0
  // ProsodyAligner (lse.py:59-81), forcing=False
  if (!c.dry) SSB_CUDA(cudaMemcpyAsync(style, dec0, (size_t)sf.rows * H * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  __half *sth = nullptr, *stl = nullptr, *zlh = nullptr, *zll = nullptr, *hdh = nullptr, *hdl = nullptr;
  if (tc) {
    sth = alloc_half_rows(c, sf, H); stl = alloc_half_rows(c, sf, H);
    zlh = alloc_half_rows(c, sr, H); zll = alloc_half_rows(c, sr, H);
    hdh = alloc_half_rows(c, sf, 2048); hdl = alloc_half_rows(c, sf, 2048);
    WS_OK(c);
    RUN(split_planes(c, zl, H, sr.rows, H, 1.0f, zlh, zll));
  }
  const bool atc = tc && attention_tc_enabled();
  const int64_t ldvt = (sr.rows + 7) & ~int64_t(7);
  __half *aqh = nullptr, *aql = nullptr, *akh = nullptr, *akl = nullptr, *avth = nullptr, *avtl = nullptr;
  if (atc) {
    aqh = alloc_half_rows(c, sf, H); aql = alloc_half_rows(c, sf, H);
    akh = alloc_half_rows(c, sr, 2 * H); akl = alloc_half_rows(c, sr, 2 * H);
    avth = c.alloc<__half>((size_t)ldvt * H); avtl = c.alloc<__half>((size_t)ldvt * H);
    WS_OK(c);
  }
  auto tcg = [&](const SeqDev& sq, const __half* ah, const __half* al, const ConvTC& w) {
    GemmTC g;
    g.A_hi = ah; g.A_lo = al; g.rows_total = sq.rows; g.w = &w; g.tiles = sq.tiles; g.ntiles = sq.ntiles;
    g.e.mode = EPI_GENERIC;
    return g;
  };
  for (int i = 0; i < 2; ++i) {
    const AlignLayer& L = m.align[i];
    if (tc) {
      RUN(split_planes(c, style, H, sf.rows, H, 1.0f, sth, stl));
      GemmTC gq = tcg(sf, sth, stl, L.q_tc);
      if (atc) { gq.e.oh = aqh; gq.e.ol = aql; gq.e.ldh = H; }
      else { gq.e.out = q; gq.e.ldo = H; }
      RUN(conv_gemm_tc(c, gq));
      GemmTC gk = tcg(sr, zlh, zll, L.kv_tc);
      if (atc) { gk.e.oh = akh; gk.e.ol = akl; gk.e.ldh = 2 * H; }
      else { gk.e.out = kv; gk.e.ldo = 2 * H; }
      RUN(conv_gemm_tc(c, gk));
    } else {
      {
        ConvGemm g = make_gemm(L.q, sf, style, H);
        g.e.out = q; g.e.ldo = H;
        RUN(conv_gemm(c, g));
      }
      {
        ConvGemm g = make_gemm(L.kv, sr, zl, H);
        g.e.out = kv; g.e.ldo = 2 * H;
        RUN(conv_gemm(c, g));
      }
    }
    if (atc) {  // cross-attention on the tcgen05 kernel: q planes [F rows, 256], k | v planes [R rows, 512]
      RUN(transpose_planes(c, akh, akl, 2 * H, H, sr.rows, H, avth, avtl, ldvt));
      AttnTCArgs a;
      a.utt_q = sf.utt; a.utt_k = sr.utt; a.B = sf.B; a.max_q = sf.maxlen; a.heads = 2;
      a.Qh = aqh; a.Ql = aql; a.rows_q = sf.rows; a.ldq = H; a.qcol0 = 0;
      a.Kh = akh; a.Kl = akl; a.rows_k = sr.rows; a.ldk = 2 * H; a.kcol0 = 0;
      a.Vth = avth; a.Vtl = avtl; a.ldvt = ldvt;
      a.keymask = kmask; a.scale = 0.08838834764831845f;
      a.oh = sth; a.ol = stl; a.ldh = H;
      RUN(attention_tc(c, a));
    } else {
      AttnArgs a;
      a.utt_q = sf.utt; a.utt_k = sr.utt; a.B = sf.B; a.max_q = sf.maxlen; a.heads = 2;
      a.Q = q; a.ldq = H; a.K = kv; a.ldk = 2 * H; a.V = kv + H; a.ldv = 2 * H;
      a.keymask = kmask; a.scale = 0.08838834764831845f; a.out = att; a.ldo = H;
      RUN(attention(c, a));
    }
    if (tc) {
      if (!atc) RUN(split_planes(c, att, H, sf.rows, H, 1.0f, sth, stl));
      GemmTC g = tcg(sf, sth, stl, L.out_tc);
      g.e.res = style; g.e.ld_res = H; g.e.out = tmp; g.e.ldo = H;
      RUN(conv_gemm_tc(c, g));
    } else {
      ConvGemm g = make_gemm(L.out, sf, att, H);
      g.e.res = style; g.e.ld_res = H; g.e.out = tmp; g.e.ldo = H;
      RUN(conv_gemm(c, g));
    }
    RUN(layernorm_rows(c, sf, tmp, H, style, H, H, L.n1_g, L.n1_b, 1e-5f, nullptr));
    if (tc) {
      RUN(split_planes(c, style, H, sf.rows, H, 1.0f, sth, stl));
      GemmTC g1 = tcg(sf, sth, stl, L.lin1_tc);
      g1.e.act = ACT_RELU; g1.e.oh = hdh; g1.e.ol = hdl; g1.e.ldh = 2048;
      RUN(conv_gemm_tc(c, g1));
      GemmTC g2 = tcg(sf, hdh, hdl, L.lin2_tc);
      g2.e.res = style; g2.e.ld_res = H; g2.e.out = tmp; g2.e.ldo = H;
      RUN(conv_gemm_tc(c, g2));
    } else {
      {
        ConvGemm g = make_gemm(L.lin1, sf, style, H);
        g.e.act = ACT_RELU; g.e.out = hid; g.e.ldo = 2048;
        RUN(conv_gemm(c, g));
      }
      {
        ConvGemm g = make_gemm(L.lin2, sf, hid, 2048);
        g.e.res = style; g.e.ld_res = H; g.e.out = tmp; g.e.ldo = H;
        RUN(conv_gemm(c, g));
      }
    }
    RUN(layernorm_rows(c, sf, tmp, H, style, H, H, L.n2_g, L.n2_b, 1e-5f, nullptr));
  }
  c.release(mk);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// a13/a18: the denoiser residual stack for one diffusion step.
// On entry b.x holds the residual stream and (b.y | b.yh,b.yl) holds x + d[t][0]; both are overwritten.
// SIMT path: fp32 conv_gemm.  Tensor-core path (b.tc): tcgen05 GEMMs on fp16 hi/lo planes.
// The two tcgen05 GEMMs of residual layer l over the row tiles [tiles, tiles + ntiles) (net.py:66-78)
static GemmTC layer_gate_gemm(const Denoiser& d, const DenoiserBufs& b, int64_t rows, const int2* tiles, int ntiles, int l) {
  const int C = d.C;
  GemmTC g;
  g.A_hi = b.yh; g.A_lo = b.yl; g.rows_total = rows; g.w = &d.layers[l].dil_tc; g.tiles = tiles; g.ntiles = ntiles;
  if (b.condpre) {  // conditioner hoisted: K = 3*C only, the projection arrives as an epilogue addend (one [rows, 2C] matrix per layer)
    g.e.add = b.condpre + (size_t)l * (size_t)rows * 2 * C; g.e.ld_add = 2 * C;
  } else {
    g.A2_hi = b.ch; g.A2_lo = b.cl; g.w2 = &d.layers[l].cond_tc;  // K = 3*C (taps of y) + 256 (cond)
  }
  g.e.mode = EPI_GATE; g.e.bias = d.layers[l].bias_gate_tc;
  g.e.oh = b.zh; g.e.ol = b.zl; g.e.ldh = C;
  return g;
}
static GemmTC layer_res_gemm(const Denoiser& d, const DenoiserBufs& b, int64_t rows, const int2* tiles, int ntiles, int l,
                             const float* dt, int tile0 = 0 /* index of tiles[0] in the layout's tile table */) {
  const int C = d.C, L = d.L;
  GemmTC g;
  g.A_hi = b.zh; g.A_lo = b.zl; g.rows_total = rows; g.w = &d.layers[l].outp_tc; g.tiles = tiles; g.ntiles = ntiles;
  // residual stream carried ONLY as the fp16 hi/lo planes of y = x + step bias (in place: this epilogue reads
  // y_l[row] and writes y_{l+1}[row] for the same rows/columns): no fp32 x is read or written in the T x L loop
  g.e.mode = EPI_RES_SKIP; g.e.C = C; g.e.beta = 0.70710678118654752440f;
  g.e.rh = b.yh; g.e.rl = b.yl; g.e.ld_rh = C; g.e.vec1 = dt + (size_t)l * C;
  if (l + 1 < L) { g.e.oh = b.yh; g.e.ol = b.yl; g.e.ldh = C; g.e.vec2 = dt + (size_t)(l + 1) * C; }
  g.e.skip = b.skip; g.e.ld_skip = C; g.e.skip_init = (l == 0);
  g.e.skip_tiled = b.skip_tiled ? 1 : 0; g.e.tile_base = tile0;
  if (b.tc_heads && l == L - 1) { g.e.sh = b.skh; g.e.sl = b.skl; }
  return g;
}

// One independent sub-problem of a sampler step: a denoiser, its buffers and the row tiles it covers.
struct Lane {
  const Denoiser* d;
  const DenoiserBufs* b;
  int64_t rows;
  const int2* tiles;
  int ntiles;
  int t;
  int tile0;  // index of tiles[0] in the layout's tile table
};
// Residual layers of TWO independent lanes, software-pipelined by half a layer: every launch but the first and the last is
// the interleaved dual kernel (conv_gemm_tc_dual) working on the gate conv of one lane and the 1x1 residual conv of the
// other, so the MMA-bound and the HBM-bound half of a layer overlap inside every SM.  2L + 1 launches instead of 2 x 2L.
static int denoiser_layers_dual(Ctx& c, const Lane& A, const Lane& B) {
  const int L = A.d->L;
  const float* dtA = A.d->dtab + (size_t)A.t * L * A.d->C;
  const float* dtB = B.d->dtab + (size_t)B.t * L * B.d->C;
  RUN(conv_gemm_tc(c, layer_gate_gemm(*A.d, *A.b, A.rows, A.tiles, A.ntiles, 0)));
  for (int l = 0; l < L; ++l) {
    RUN(conv_gemm_tc_dual(c, layer_gate_gemm(*B.d, *B.b, B.rows, B.tiles, B.ntiles, l),
                          layer_res_gemm(*A.d, *A.b, A.rows, A.tiles, A.ntiles, l, dtA, A.tile0)));
    if (l + 1 < L)
      RUN(conv_gemm_tc_dual(c, layer_gate_gemm(*A.d, *A.b, A.rows, A.tiles, A.ntiles, l + 1),
                            layer_res_gemm(*B.d, *B.b, B.rows, B.tiles, B.ntiles, l, dtB, B.tile0)));
    else
      RUN(conv_gemm_tc(c, layer_res_gemm(*B.d, *B.b, B.rows, B.tiles, B.ntiles, l, dtB, B.tile0)));
  }
  return 0;
}
static bool dual_lane_ok(const Denoiser& d, const DenoiserBufs& b) {
  return b.tc && b.condpre != nullptr && dual_enabled() && d.layers[0].dil_tc.hb == d.layers[0].outp_tc.hb &&
         (d.layers[0].dil_tc.hb == 128 || d.layers[0].dil_tc.hb == 96);
}

// skip_projection + output_projection of the denoiser (net.py:126-130 / :262-266)
static int denoiser_heads(Ctx& c, const Denoiser& d, const SeqDev& s, DenoiserBufs& b) {
  const int C = d.C, L = d.L;
  if (b.tc_heads) {
    {  // skip_projection (1/sqrt(L) folded into the packed weights) + ReLU -> planes
      GemmTC g;
      g.A_hi = b.skh; g.A_lo = b.skl; g.rows_total = s.rows; g.w = &d.skip_tc; g.tiles = s.tiles; g.ntiles = s.ntiles;
      g.e.mode = EPI_GENERIC; g.e.bias = d.skip_bias_pad; g.e.act = ACT_RELU; g.e.oh = b.sh; g.e.ol = b.sl; g.e.ldh = C;
      g.e.n_valid = C;
      RUN(conv_gemm_tc(c, g));
    }
    {  // output_projection (N padded to a tile multiple; only the first out_dims columns are meaningful)
      GemmTC g;
      g.A_hi = b.sh; g.A_lo = b.sl; g.rows_total = s.rows; g.w = &d.out_tc; g.tiles = s.tiles; g.ntiles = s.ntiles;
      g.e.mode = EPI_GENERIC; g.e.bias = d.out_bias_pad; g.e.out = b.head; g.e.ldo = b.ld_head;
      g.e.n_valid = (d.out_dims + 31) / 32 * 32;
      RUN(conv_gemm_tc(c, g));
    }
    return 0;
  }
  {
    ConvGemm g = make_gemm(d.skip_proj, s, b.skip, C);
    g.a_scale = 1.0f / sqrtf((float)L);
    g.e.act = ACT_RELU; g.e.out = b.sbuf; g.e.ldo = C;
    RUN(conv_gemm(c, g));
  }
  {
    ConvGemm g = make_gemm(d.out_proj, s, b.sbuf, C);
    g.e.out = b.head; g.e.ldo = b.ld_head;
    RUN(conv_gemm(c, g));
  }
  return 0;
}

// `split` (0 = off): row tiles [0, split) and [split, ntiles) belong to two disjoint sets of utterances; the residual layers
// then run as two software-pipelined lanes on the interleaved dual kernel (denoiser_layers_dual).
int denoiser_stack(Ctx& c, const Denoiser& d, const SeqDev& s, int t, DenoiserBufs& b, int split) {
  const int C = d.C, L = d.L;
  SSB_CHECK(d.dtab != nullptr && t >= 0 && t < d.T, "denoiser: schedule not set (ssb_model_set_schedule) or bad t");
  const float* dt = d.dtab + (size_t)t * L * C;
  if (split > 0 && split < s.ntiles && dual_lane_ok(d, b) && !c.dry) {
    const Lane A{&d, &b, s.rows, s.tiles, split, t, 0}, B{&d, &b, s.rows, s.tiles + split, s.ntiles - split, t, split};
    RUN(denoiser_layers_dual(c, A, B));
    return denoiser_heads(c, d, s, b);
  }
  for (int l = 0; l < L; ++l) {
    if (b.tc) {
      RUN(conv_gemm_tc(c, layer_gate_gemm(d, b, s.rows, s.tiles, s.ntiles, l)));
      RUN(conv_gemm_tc(c, layer_res_gemm(d, b, s.rows, s.tiles, s.ntiles, l, dt)));
      continue;
    }
    {
      ConvGemm g = make_gemm(d.layers[l].dil, s, b.y, C);
      g.e.mode = EPI_GATE; g.e.add = b.condall + (size_t)l * 2 * C; g.e.ld_add = L * 2 * C; g.e.out = b.zg; g.e.ldo = C;
      RUN(conv_gemm(c, g));
    }
    {
      ConvGemm g = make_gemm(d.layers[l].outp, s, b.zg, C);
      g.e.mode = EPI_RES_SKIP; g.e.C = C; g.e.res = b.x; g.e.ld_res = C; g.e.beta = 0.70710678118654752440f;
      g.e.out = b.x; g.e.ldo = C;
      if (l + 1 < L) { g.e.out2 = b.y; g.e.ldo2 = C; g.e.vec2 = dt + (size_t)(l + 1) * C; }
      g.e.skip = b.skip; g.e.ld_skip = C; g.e.skip_init = (l == 0);
      RUN(conv_gemm(c, g));
    }
  }
  return denoiser_heads(c, d, s, b);
}

static __half* alloc_half_rows(Ctx& c, const SeqDev& s, int C) {
  __half* p = c.alloc<__half>((size_t)s.rows * C);
  if (!c.dry && p && !c.failed) cudaMemsetAsync(p, 0, (size_t)s.rows * C * sizeof(__half), c.stream);
  return p;
}
bool denoiser_tc_ok(const Model& m, const Denoiser& d) {
  if (!m.use_tc) return false;
  for (auto& l : d.layers)
    if (!l.dil_tc.ok || !l.outp_tc.ok || !l.cond_tc.ok) return false;
  return true;
}
int alloc_denoiser(Ctx& c, const Denoiser& d, const SeqDev& s, bool tc, DenoiserBufs* b, bool hoist) {
  b->tc = tc;
  b->condpre = nullptr;
  b->x = alloc_rows(c, s, d.C);
  b->y = b->zg = nullptr;
  b->yh = b->yl = b->zh = b->zl = b->ch = b->cl = nullptr;
  b->skh = b->skl = b->sh = b->sl = b->x80h = b->x80l = nullptr;
  b->condall = nullptr;
  b->tc_heads = tc && d.skip_tc.ok && d.out_tc.ok;
  if (b->tc_heads) {
    b->skh = alloc_half_rows(c, s, d.C);
    b->skl = alloc_half_rows(c, s, d.C);
    b->sh = alloc_half_rows(c, s, d.C);
    b->sl = alloc_half_rows(c, s, d.C);
    if (!d.ddiff && d.in_tc.ok) {
      b->x80h = alloc_half_rows(c, s, 128);
      b->x80l = alloc_half_rows(c, s, 128);
    }
  }
  if (tc) {
    b->yh = alloc_half_rows(c, s, d.C);
    b->yl = alloc_half_rows(c, s, d.C);
    b->zh = alloc_half_rows(c, s, d.C);
    b->zl = alloc_half_rows(c, s, d.C);
    b->ch = alloc_half_rows(c, s, 256);
    b->cl = alloc_half_rows(c, s, 256);
    // only valid rows are ever written or read (epilogues are row-bounded): no 4.6 GB memset for batch64
    if (hoist && d.cond_all_tc.ok) b->condpre = alloc_rows(c, s, d.L * 2 * d.C, false);
  } else {
    b->y = alloc_rows(c, s, d.C);
    b->zg = alloc_rows(c, s, d.C);
    b->condall = alloc_rows(c, s, d.L * 2 * d.C, false);
  }
  // tensor-core heads: the fp32 skip accumulator is private to the RES_SKIP epilogue (the heads read the planes the last
  // layer writes), so it is kept chunk-tiled (every 32 x 32 epilogue chunk one contiguous 4 KB block; conv_gemm_tc.cuh)
  static const bool skip_rowmajor = getenv("SSB_SKIP_ROWMAJOR") != nullptr;  // A/B switch for the layout experiment
  b->skip_tiled = b->tc_heads && !skip_rowmajor;
  if (b->skip_tiled) {
    const size_t n = (size_t)s.ntiles * TILE_M * d.C;
    b->skip = c.alloc<float>(n);  // fully written by layer 0 (skip_init) before it is read: no memset
  } else {
    b->skip = alloc_rows(c, s, d.C);
  }
  b->sbuf = b->tc_heads ? nullptr : alloc_rows(c, s, d.C);
  b->ld_head = b->tc_heads ? d.out_tc.N : ((d.out_dims + 3) & ~3);
  b->head = alloc_rows(c, s, b->ld_head);
  WS_OK(c);
  return 0;
}
// Conditioner: tensor-core path -> fp16 hi/lo planes of cond (contracted inside every layer GEMM);
// SIMT path -> the step-invariant projection of all L layers hoisted into one [rows, L*2C] buffer.
int prepare_cond(Ctx& c, const Denoiser& d, const SeqDev& s, const float* cond_g, DenoiserBufs& b) {
  if (b.tc) {
    RUN(split_planes(c, cond_g, 256, s.rows, 256, 1.0f, b.ch, b.cl));
    if (!b.condpre) return 0;
    GemmTC g;  // all L conditioner projections at once: [rows, 256] x [256, L*2C], once per sampler call
    g.A_hi = b.ch; g.A_lo = b.cl; g.rows_total = s.rows; g.w = &d.cond_all_tc; g.tiles = s.tiles; g.ntiles = s.ntiles;
    g.e.mode = EPI_GENERIC; g.e.out = b.condpre; g.e.ldo = d.L * 2 * d.C;
    g.e.out_nb = 2 * d.C; g.e.out_bs = (int64_t)s.rows * 2 * d.C;  // layer-major: row pitch 2C floats instead of L * 2C
    return conv_gemm_tc(c, g);
  }
  ConvGemm g = make_gemm(d.cond_all, s, cond_g, 256);
  g.e.out = b.condall; g.e.ldo = d.L * 2 * d.C;
  return conv_gemm(c, g);
}

// a18 entry for one evaluation (mel): x80 [rows,80] guarded -> head
int lane_split_tile(const Seq* q) {
  if (!q || q->B < 2) return 0;
  const int total = q->ntiles(1);
  int acc = 0, best = 0;
  for (int b = 0; b + 1 < q->B; ++b) {
    acc += (q->len[b] + TILE_M - 1) / TILE_M;
    if (best == 0 || abs(2 * acc - total) < abs(2 * best - total)) best = acc;
  }
  // a lopsided split leaves most tiles without a partner of the other kind: not worth the extra launch
  return (best * 4 >= total && best * 4 <= 3 * total) ? best : 0;
}

int mel_denoiser_eval(Ctx& c, const Denoiser& d, const SeqDev& s, int t, const float* x80, DenoiserBufs& b, int split) {
  if (b.x80h) {  // tensor-core input projection: K padded 80 -> 128
    RUN(x80_planes(c, x80, s.rows, b.x80h, b.x80l));
    GemmTC g;
    g.A_hi = b.x80h; g.A_lo = b.x80l; g.rows_total = s.rows; g.w = &d.in_tc; g.tiles = s.tiles; g.ntiles = s.ntiles;
    g.e.mode = EPI_GENERIC; g.e.bias = d.in_proj.bias; g.e.act = ACT_RELU;  // planes of y = relu(in_proj) + step bias only
    g.e.oh = b.yh; g.e.ol = b.yl; g.e.ldh = d.C; g.e.vec2 = d.dtab + (size_t)t * d.L * d.C;
    RUN(conv_gemm_tc(c, g));
    return denoiser_stack(c, d, s, t, b, split);
  }
  ConvGemm g = make_gemm(d.in_proj, s, x80, 80);
  g.e.act = ACT_RELU; g.e.out = b.x; g.e.ldo = d.C;
  g.e.vec2 = d.dtab + (size_t)t * d.L * d.C;
  if (b.tc) { g.e.out2_h = b.yh; g.e.out2_l = b.yl; g.e.ldh = d.C; }
  else { g.e.out2 = b.y; g.e.ldo2 = d.C; }
  RUN(conv_gemm(c, g));
  return denoiser_stack(c, d, s, t, b, split);
}

// a18+a19, single launch: all T reverse steps in the persistent tcgen05 kernel (sampler_tc.cu).
static int run_mel_diffusion_persistent(Ctx& c, const Model& m, const SeqDev& s, const float* cond_g, const float* coarse_g,
                                        const float* noise, uint64_t seed, float* mel_tight) {
  const Denoiser& d = m.melnet;
  const int C = d.C, L = d.L, T = d.T;
  const int CS = 4;  // cluster size along N: A tiles are TMA-multicast to the 4 CTAs that share an M-tile
  const size_t mk = c.mark();
  float* xm = alloc_rows(c, s, 80);
  float* x = alloc_rows(c, s, C);
  float* skip = alloc_rows(c, s, C);
  __half* pl[12];
  const int pcols[6] = {128, C, C, 256, C, C};  // x80, y, z, cond, skip, s
  for (int i = 0; i < 6; ++i) {
    pl[2 * i] = alloc_half_rows(c, s, pcols[i]);
    pl[2 * i + 1] = alloc_half_rows(c, s, pcols[i]);
  }
  const int nmaps = 12 + 2 + 6 * L + 4;
  const int nph = T * (2 * L + 3);
  CUtensorMap* maps_dev = c.alloc<CUtensorMap>((size_t)nmaps);
  SPhase* ph_dev = c.alloc<SPhase>((size_t)nph);
  unsigned* ctr = c.alloc<unsigned>(4);
  WS_OK(c);
  const size_t per = (size_t)s.total * 80;
  const float sa = c.dry ? 0.f : d.gtab_h[(size_t)(T - 1) * 8 + 5], s1a = c.dry ? 0.f : d.gtab_h[(size_t)(T - 1) * 8 + 6];
  RUN(mel_q_sample(c, s, coarse_g, 80, noise, m.spec_min, m.spec_max, sa, s1a, xm, 80, seed, 1000));
  RUN(x80_planes(c, xm, s.rows, pl[0], pl[1]));
  RUN(split_planes(c, cond_g, 256, s.rows, 256, 1.0f, pl[6], pl[7]));
  // step-invariant conditioner projection of all L layers, once per call (one per-launch GEMM): the gate phases then
  // contract K = 3C instead of 3C + 256 and add it in the epilogue
  float* condpre = nullptr;
  if (m.cond_hoist && d.cond_all_tc.ok) {
    condpre = alloc_rows(c, s, L * 2 * C, false);
    WS_OK(c);
    GemmTC g;
    g.A_hi = pl[6]; g.A_lo = pl[7]; g.rows_total = s.rows; g.w = &d.cond_all_tc; g.tiles = s.tiles; g.ntiles = s.ntiles;
    g.e.mode = EPI_GENERIC; g.e.out = condpre; g.e.ldo = L * 2 * C;
    g.e.out_nb = 2 * C; g.e.out_bs = (int64_t)s.rows * 2 * C;
    RUN(conv_gemm_tc(c, g));
  }
  if (!c.dry) {
    std::vector<CUtensorMap> maps((size_t)nmaps);
    for (int i = 0; i < 6; ++i) {
      if (make_act_map(&maps[2 * i], pl[2 * i], s.rows, pcols[i], 128 / CS)) return -1;
      if (make_act_map(&maps[2 * i + 1], pl[2 * i + 1], s.rows, pcols[i], 128 / CS)) return -1;
    }
    auto put = [&](int idx, const ConvTC& w) { maps[idx] = w.tm_hi[1]; maps[idx + 1] = w.tm_lo[1]; };
    const int W_IN = 12, W_L0 = 14, W_SKIP = 14 + 6 * L, W_OUT = W_SKIP + 2;
    put(W_IN, d.in_tc);
    for (int l = 0; l < L; ++l) {
      put(W_L0 + 6 * l, d.layers[l].dil_tc);
      put(W_L0 + 6 * l + 2, d.layers[l].cond_tc);
      put(W_L0 + 6 * l + 4, d.layers[l].outp_tc);
    }
    put(W_SKIP, d.skip_tc);
    put(W_OUT, d.out_tc);
    std::vector<SPhase> ph((size_t)nph);
    size_t k = 0;
    for (int t = T - 1; t >= 0; --t) {
      const float* dt = d.dtab + (size_t)t * L * C;
      SPhase z;
      memset(&z, 0, sizeof(z));
      z.a2 = -1; z.taps = 1; z.dil = 1; z.beta = 1.0f; z.sync_after = 1;
      {  // input_projection + ReLU ; y = x + step bias of layer 0
        SPhase q = z;
        q.a1 = 0; q.w1 = W_IN; q.kchunks = 2; q.N = C; q.NT = C / 64; q.mode = SP_INPROJ; q.bias = d.in_proj.bias;
        q.out = x; q.ldo = C; q.oh = pl[2]; q.ol = pl[3]; q.ldh = C; q.vec2 = dt;
        ph[k++] = q;
      }
      for (int l = 0; l < L; ++l) {
        SPhase a = z;  // dilated conv (3 taps of y) + conditioner (cond) -> gate -> z planes
        a.a1 = 2; a.a2 = 6; a.w1 = W_L0 + 6 * l; a.w2 = W_L0 + 6 * l + 2; a.taps = 3; a.kchunks = C / 64; a.kchunks2 = 4;
        a.dil = d.layers[l].dil_tc.dil; a.center = 1; a.N = 2 * C; a.NT = 2 * C / 64; a.mode = SP_GATE;
        a.bias = d.layers[l].bias_gate_tc; a.oh = pl[4]; a.ol = pl[5]; a.ldh = C;
        if (condpre) { a.a2 = -1; a.kchunks2 = 0; a.add = condpre + (size_t)l * (size_t)s.rows * 2 * C; a.ld_add = 2 * C; }
        ph[k++] = a;
        SPhase b = z;  // 1x1 output projection -> residual stream, next layer's input planes, skip sum
        b.a1 = 4; b.w1 = W_L0 + 6 * l + 4; b.kchunks = C / 64; b.N = 2 * C; b.NT = 2 * C / 64; b.mode = SP_RES_SKIP;
        b.bias = d.layers[l].outp.bias; b.res = x; b.ld_res = C; b.out = x; b.ldo = C; b.beta = 0.70710678118654752440f;
        if (l + 1 < L) { b.oh = pl[2]; b.ol = pl[3]; b.ldh = C; b.vec2 = dt + (size_t)(l + 1) * C; }
        b.skip = skip; b.ld_skip = C; b.C = C; b.skip_init = (l == 0);
        if (l == L - 1) { b.sh = pl[8]; b.sl = pl[9]; }
        ph[k++] = b;
      }
      {  // skip_projection (1/sqrt(L) folded into the weights) + ReLU -> s planes
        SPhase q = z;
        q.a1 = 8; q.w1 = W_SKIP; q.kchunks = C / 64; q.N = C; q.NT = C / 64; q.mode = SP_SKIPPROJ; q.bias = d.skip_proj.bias;
        q.oh = pl[10]; q.ol = pl[11]; q.ldh = C;
        ph[k++] = q;
      }
      {  // output_projection -> eps ; fused DDPM posterior step on x_t
        SPhase q = z;
        q.a1 = 10; q.w1 = W_OUT; q.kchunks = C / 64; q.N = 256; q.NT = 4; q.mode = SP_MEL_SAMPLE; q.bias = d.out_bias_pad;
        q.out = xm; q.ldo = 80; q.oh = pl[0]; q.ol = pl[1]; q.ldh = 128; q.tab = d.gtab + (size_t)t * 8;
        q.noise = noise ? noise + per * (size_t)(T - t) : nullptr; q.seed = seed; q.stream_id = 1001 + (uint64_t)t; q.n_valid = 80;
        ph[k++] = q;
      }
    }
    SSB_CUDA(cudaMemcpyAsync(maps_dev, maps.data(), sizeof(CUtensorMap) * nmaps, cudaMemcpyHostToDevice, c.stream));
    SSB_CUDA(cudaMemcpyAsync(ph_dev, ph.data(), sizeof(SPhase) * nph, cudaMemcpyHostToDevice, c.stream));
    RUN(launch_sampler_tc(c, maps_dev, ph_dev, nph, s.tiles, s.tile_tight, s.ntiles, 2 * C / 64, ctr, CS));
  }
  RUN(mel_denorm(c, s, xm, 80, m.spec_min, m.spec_max, nullptr, mel_tight, 80));
  c.release(mk);
  return 0;
}

// a18+a19: DiffusionDecoder.forward(infer=True) (shallow_diffusion_tts.py:284-307)
int run_mel_diffusion(Ctx& c, const Model& m, const SeqDev& s, const float* cond_g, const float* coarse_g,
                      const float* noise /*tight [(T+1), total, 80] or null*/, uint64_t seed, float* mel_tight,
                      const Seq* host_seq) {
  const Denoiser& d = m.melnet;
  SSB_CHECK(d.T > 0, "mel schedule not set: call ssb_model_set_schedule(which=0)");
  // Utterances are independent, so the T x L loop can run per GROUP of consecutive utterances (a contiguous slice of the
  // guard-banded layout: the sub-batch simply aliases the big buffers).  Two users, production (Philox) mode only - the
  // injected-noise tensors are strided by the whole batch; each group gets its own seed:
  //  * SSB_MEL_GROUP_FRAMES=<n> (experiment): groups of ~n frames whose working set stays L2-resident;
  //  * ssb_model_set_persistent_groups(1): groups of <= 48 row tiles, each run by the single-launch persistent kernel
  //    (BASELINE.json configs[4]: persistent-kernel vs per-step-launch at batch 64).
  if (host_seq && !noise && !c.dry) {
    const char* ge = getenv("SSB_MEL_GROUP_FRAMES");
    const long gf = ge ? atol(ge) : 0;
    const bool by_tiles = m.persistent_groups && m.persistent && s.ntiles > 48;
    if (by_tiles || (gf > 0 && host_seq->total > gf + gf / 2)) {
      int b0 = 0;
      int64_t tight0 = 0;
      int gi = 0;
      auto tiles_of = [&](int b) { return (host_seq->len[b] + TILE_M - 1) / TILE_M; };
      while (b0 < host_seq->B) {
        int b1 = b0;
        int64_t fr = 0, nt = 0;
        while (b1 < host_seq->B &&
               (b1 == b0 || (by_tiles ? nt + tiles_of(b1) <= 48 : fr + host_seq->len[b1] <= gf))) {
          nt += tiles_of(b1);
          fr += host_seq->len[b1++];
        }
        std::vector<int32_t> offs((size_t)(b1 - b0) + 1, 0);
        for (int b = b0; b < b1; ++b) offs[(size_t)(b - b0) + 1] = offs[(size_t)(b - b0)] + host_seq->len[b];
        Seq q;
        q.build(offs.data(), b1 - b0);
        const size_t mkg = c.mark();
        SeqDev sg;
        RUN(upload_layout(c, q, 1, &sg));
        const int64_t row_off = (int64_t)host_seq->rs[b0] - GUARD;  // the sub-layout's row 0 inside the big buffers
        RUN(run_mel_diffusion(c, m, sg, cond_g + row_off * 256, coarse_g + row_off * 80, nullptr,
                              seed + 0x9E3779B97F4A7C15ull * (uint64_t)gi, mel_tight + tight0 * 80, nullptr));
        c.release(mkg);
        tight0 += fr;
        b0 = b1;
        ++gi;
      }
      return 0;
    }
  }
  if (m.persistent && denoiser_tc_ok(m, d) && d.in_tc.ok && d.skip_tc.ok && d.out_tc.ok && s.ntiles <= 48 &&
      sampler_tc_max_ctas() > 0)
    return run_mel_diffusion_persistent(c, m, s, cond_g, coarse_g, noise, seed, mel_tight);
  const size_t mk = c.mark();
  DenoiserBufs b;
  RUN(alloc_denoiser(c, d, s, denoiser_tc_ok(m, d), &b, m.cond_hoist));
  float* xm = alloc_rows(c, s, 80);
  WS_OK(c);
  RUN(prepare_cond(c, d, s, cond_g, b));
  const int T = d.T;
  const size_t per = (size_t)s.total * 80;
  const float sa = c.dry ? 0.f : d.gtab_h[(size_t)(T - 1) * 8 + 5], s1a = c.dry ? 0.f : d.gtab_h[(size_t)(T - 1) * 8 + 6];
  RUN(mel_q_sample(c, s, coarse_g, 80, noise, m.spec_min, m.spec_max, sa, s1a, xm, 80, seed, 1000));
  // two utterance groups -> the residual layers run as two lanes on the interleaved dual kernel
  const int split = s.ntiles >= 148 ? lane_split_tile(host_seq) : 0;
  for (int t = T - 1; t >= 0; --t) {
    RUN(mel_denoiser_eval(c, d, s, t, xm, b, split));
    const float* nz = noise ? noise + per * (size_t)(T - t) : nullptr;
    RUN(mel_p_sample(c, s, xm, 80, b.head, b.ld_head, nz, d.gtab + (size_t)t * 8, seed, 1001 + (uint64_t)t));
  }
  RUN(mel_denorm(c, s, xm, 80, m.spec_min, m.spec_max, nullptr, mel_tight, 80));
  c.release(mk);
  return 0;
}

// f2: PLMS / PNDM sampler over the same denoiser (GaussianDiffusion.p_sample_plms + the pndm_speedup loop of
// GaussianDiffusion.forward, shallow_diffusion_tts.py:164-197,254-260): T / interval evaluations (+1 for the first step).
int run_mel_diffusion_plms(Ctx& c, const Model& m, const SeqDev& s, const float* cond_g, const float* coarse_g,
                           const float* q_noise /*tight [total, 80] or null*/, uint64_t seed, int interval, float* mel_tight) {
  const Denoiser& d = m.melnet;
  SSB_CHECK(d.T > 0, "mel schedule not set: call ssb_model_set_schedule(which=0)");
  SSB_CHECK(interval >= 1 && interval < d.T, "plms: interval (pndm_speedup) must be in [1, T)");
  const size_t mk = c.mark();
  DenoiserBufs b;
  RUN(alloc_denoiser(c, d, s, denoiser_tc_ok(m, d), &b, m.cond_hoist));
  float* xm = alloc_rows(c, s, 80);
  float* xp = alloc_rows(c, s, 80);
  float* hist[3] = {alloc_rows(c, s, 80), alloc_rows(c, s, 80), alloc_rows(c, s, 80)};
  WS_OK(c);
  RUN(prepare_cond(c, d, s, cond_g, b));
  const int T = d.T;
  auto acp = [&](int t) { return c.dry ? 0.5f : d.gtab_h[(size_t)t * 8 + 7]; };
  const float sa = c.dry ? 0.f : d.gtab_h[(size_t)(T - 1) * 8 + 5], s1a = c.dry ? 0.f : d.gtab_h[(size_t)(T - 1) * 8 + 6];
  RUN(mel_q_sample(c, s, coarse_g, 80, q_noise, m.spec_min, m.spec_max, sa, s1a, xm, 80, seed, 1000));
  int nh = 0;  // predictions in the history; hist[(head + k) % 3] is the k-th newest
  int head = 0;
  int t0 = 0;
  for (int t = 0; t < T; t += interval) t0 = t;  // reversed(range(0, T, interval)) starts at the largest multiple
  for (int t = t0; t >= 0; t -= interval) {
    const int tp = t - interval > 0 ? t - interval : 0;
    RUN(mel_denoiser_eval(c, d, s, t, xm, b));
    PlmsArgs a;
    a.x = xm; a.eps = b.head; a.lde = b.ld_head; a.a_t = acp(t); a.a_prev = acp(tp);
    const int slot = (head + 2) % 3;  // overwritten by this step's eps: the oldest entry
    if (nh == 0) {
      // second-order start: trial step with eps, second evaluation at the trial point, average (:182-185)
      PlmsArgs p1 = a;
      p1.x_out = xp; p1.hist_out = hist[slot];  // eps is saved before the head buffer is overwritten by the 2nd evaluation
      RUN(plms_update(c, s, p1));
      RUN(mel_denoiser_eval(c, d, s, tp, xp, b));
      a.h1 = hist[slot]; a.w0 = 1.f; a.w1 = 1.f; a.den = 2.f;  // (eps + eps_prev) / 2 with eps_prev = current head
      a.x_out = xm; a.hist_out = nullptr;
      RUN(plms_update(c, s, a));
    } else {
      const float* h1 = hist[head];
      const float* h2 = hist[(head + 1) % 3];
      const float* h3 = hist[(head + 2) % 3];
      if (nh == 1) { a.h1 = h1; a.w0 = 3.f; a.w1 = -1.f; a.den = 2.f; }
      else if (nh == 2) { a.h1 = h1; a.h2 = h2; a.w0 = 23.f; a.w1 = -16.f; a.w2 = 5.f; a.den = 12.f; }
      else { a.h1 = h1; a.h2 = h2; a.h3 = h3; a.w0 = 55.f; a.w1 = -59.f; a.w2 = 37.f; a.w3 = -9.f; a.den = 24.f; }
      a.x_out = xm;
      // the oldest entry (h3's slot) receives eps; with nh >= 3 the same element is read (h3) and then written by one thread
      a.hist_out = hist[slot];
      RUN(plms_update(c, s, a));
    }
    head = slot;
    if (nh < 3) ++nh;
  }
  RUN(mel_denorm(c, s, xm, 80, m.spec_min, m.spec_max, nullptr, mel_tight, 80));
  c.release(mk);
  return 0;
}

// a13+a14 for BOTH F0 nets in one persistent launch: the agnostic and the specific sampler are independent
// (stylesinger.py:223-225), so each phase of the table carries two entries (one per net, no barrier between them).
int run_f0_diffusion_pair_persistent(Ctx& c, const Model& m, const SeqDev& s, const float* cond0, const float* cond1,
                                     const float* lo, const float* hi, const float* const gnoise[2],
                                     const float* const unoise[2], uint64_t seed, float* const z[2], int32_t* const uv[2]) {
  const Denoiser& d0 = m.f0net[0];
  const int C = d0.C, L = d0.L, T = d0.T;
  const int CS = 2;
  const size_t mk = c.mark();
  const int NA = 10;           // activation maps per net: y, z, cond, skip, s (hi/lo)
  const int NW = 6 * L + 4;    // weight maps per net
  const int nmaps = 2 * (NA + NW);
  const int per_step = 2 * L + 2;
  const int nph = 2 * T * per_step;
  float* x[2]; float* skip[2]; __half* pl[2][10];
  const int pcols[5] = {C, C, 256, C, C};
  for (int n = 0; n < 2; ++n) {
    x[n] = alloc_rows(c, s, C);
    skip[n] = alloc_rows(c, s, C);
    for (int i = 0; i < 5; ++i) {
      pl[n][2 * i] = alloc_half_rows(c, s, pcols[i]);
      pl[n][2 * i + 1] = alloc_half_rows(c, s, pcols[i]);
    }
  }
  CUtensorMap* maps_dev = c.alloc<CUtensorMap>((size_t)nmaps);
  SPhase* ph_dev = c.alloc<SPhase>((size_t)nph);
  unsigned* ctr = c.alloc<unsigned>(4);
  WS_OK(c);
  const size_t per = (size_t)s.total;
  for (int n = 0; n < 2; ++n) {
    const Denoiser& d = m.f0net[n];
    const uint64_t sbase = 2000 + (uint64_t)n * 100000;
    RUN(f0_init(c, s, z[n], uv[n], gnoise[n], seed, sbase));
    RUN(ddiff_input(c, s, z[n], uv[n], d.in_w, d.in_b, d.uv_emb, d.dtab + (size_t)(T - 1) * L * C, x[n], nullptr, C, pl[n][0], pl[n][1]));
    RUN(split_planes(c, n == 0 ? cond0 : cond1, 256, s.rows, 256, 1.0f, pl[n][4], pl[n][5]));
  }
  float* condpre[2] = {nullptr, nullptr};  // hoisted conditioner projections (see run_mel_diffusion_persistent)
  if (m.cond_hoist && m.f0net[0].cond_all_tc.ok && m.f0net[1].cond_all_tc.ok) {
    for (int n = 0; n < 2; ++n) {
      condpre[n] = alloc_rows(c, s, L * 2 * C, false);
      WS_OK(c);
      GemmTC g;
      g.A_hi = pl[n][4]; g.A_lo = pl[n][5]; g.rows_total = s.rows; g.w = &m.f0net[n].cond_all_tc; g.tiles = s.tiles; g.ntiles = s.ntiles;
      g.e.mode = EPI_GENERIC; g.e.out = condpre[n]; g.e.ldo = L * 2 * C;
      g.e.out_nb = 2 * C; g.e.out_bs = (int64_t)s.rows * 2 * C;
      RUN(conv_gemm_tc(c, g));
    }
  }
  if (!c.dry) {
    std::vector<CUtensorMap> maps((size_t)nmaps);
    std::vector<SPhase> ph((size_t)nph);
    // number of clusters the launcher will use: needed for the tile-group rotation of the second net
    int ncl = s.ntiles * (2 * (2 * C / 64) / CS);
    {
      const int cap = sampler_tc_max_clusters(CS);
      if (ncl > cap) ncl = cap;
      if (ncl < 1) ncl = 1;
    }
    for (int n = 0; n < 2; ++n) {
      const Denoiser& d = m.f0net[n];
      const int MB = n * (NA + NW);
      for (int i = 0; i < 5; ++i) {
        if (make_act_map(&maps[MB + 2 * i], pl[n][2 * i], s.rows, pcols[i], 128 / CS)) return -1;
        if (make_act_map(&maps[MB + 2 * i + 1], pl[n][2 * i + 1], s.rows, pcols[i], 128 / CS)) return -1;
      }
      auto put = [&](int idx, const ConvTC& w) { maps[idx] = w.tm_hi[1]; maps[idx + 1] = w.tm_lo[1]; };
      const int W_L0 = MB + NA, W_SKIP = W_L0 + 6 * L, W_OUT = W_SKIP + 2;
      for (int l = 0; l < L; ++l) {
        put(W_L0 + 6 * l, d.layers[l].dil_tc);
        put(W_L0 + 6 * l + 2, d.layers[l].cond_tc);
        put(W_L0 + 6 * l + 4, d.layers[l].outp_tc);
      }
      put(W_SKIP, d.skip_tc);
      put(W_OUT, d.out_tc);
      const uint64_t sbase = 2000 + (uint64_t)n * 100000;
      for (int ti = 0; ti < T; ++ti) {
        const int t = T - 1 - ti;
        const float* dt = d.dtab + (size_t)t * L * C;
        SPhase zp;
        memset(&zp, 0, sizeof(zp));
        zp.a2 = -1; zp.taps = 1; zp.dil = 1; zp.beta = 1.0f; zp.sync_after = (n == 1);
        // entry k of step ti for net n sits at ((ti * per_step + k) * 2 + n)
        size_t k = 0;
        auto slot = [&](size_t kk) -> SPhase& { return ph[((size_t)ti * per_step + kk) * 2 + n]; };
        for (int l = 0; l < L; ++l) {
          SPhase a = zp;
          a.a1 = MB + 0; a.a2 = MB + 4; a.w1 = W_L0 + 6 * l; a.w2 = W_L0 + 6 * l + 2; a.taps = 3; a.kchunks = C / 64; a.kchunks2 = 4;
          a.dil = d.layers[l].dil_tc.dil; a.center = 1; a.N = 2 * C; a.NT = 2 * C / 64; a.mode = SP_GATE;
          a.bias = d.layers[l].bias_gate_tc; a.oh = pl[n][2]; a.ol = pl[n][3]; a.ldh = C;
          if (condpre[n]) { a.a2 = -1; a.kchunks2 = 0; a.add = condpre[n] + (size_t)l * (size_t)s.rows * 2 * C; a.ld_add = 2 * C; }
          slot(k++) = a;
          SPhase b = zp;
          b.a1 = MB + 2; b.w1 = W_L0 + 6 * l + 4; b.kchunks = C / 64; b.N = 2 * C; b.NT = 2 * C / 64; b.mode = SP_RES_SKIP;
          b.bias = d.layers[l].outp.bias; b.res = x[n]; b.ld_res = C; b.out = x[n]; b.ldo = C; b.beta = 0.70710678118654752440f;
          if (l + 1 < L) { b.oh = pl[n][0]; b.ol = pl[n][1]; b.ldh = C; b.vec2 = dt + (size_t)(l + 1) * C; }
          b.skip = skip[n]; b.ld_skip = C; b.C = C; b.skip_init = (l == 0);
          if (l == L - 1) { b.sh = pl[n][6]; b.sl = pl[n][7]; }
          slot(k++) = b;
        }
        {
          SPhase q = zp;
          q.a1 = MB + 6; q.w1 = W_SKIP; q.kchunks = C / 64; q.N = d.skip_tc.N; q.NT = d.skip_tc.N / 64; q.mode = SP_SKIPPROJ;
          q.bias = d.skip_bias_pad; q.oh = pl[n][8]; q.ol = pl[n][9]; q.ldh = C; q.n_valid = C;
          slot(k++) = q;
        }
        {
          SPhase q = zp;
          q.a1 = MB + 8; q.w1 = W_OUT; q.kchunks = C / 64; q.N = d.out_tc.N; q.NT = d.out_tc.N / 64; q.mode = SP_F0_SAMPLE;
          q.bias = d.out_bias_pad; q.out = z[n]; q.uv = uv[n]; q.clip_lo = lo; q.clip_hi = hi;
          q.tab = d.gtab + (size_t)t * 8; q.tab2 = d.mtab + (size_t)t * 8; q.tstep = t; q.log_eps = m.log_eps;
          q.noise = gnoise[n] ? gnoise[n] + per * (size_t)(T - t) : nullptr;
          q.noise2 = unoise[n] ? unoise[n] + per * 2 * (size_t)(T - 1 - t) : nullptr;
          q.seed = seed; q.stream_id = sbase + 10 + 2 * (uint64_t)t;
          q.has_next = t > 0; q.C = C; q.in_w = d.in_w; q.in_b = d.in_b; q.uv_emb = d.uv_emb; q.x_next = x[n];
          q.oh = pl[n][0]; q.ol = pl[n][1]; q.ldh = C;
          q.vec2 = t > 0 ? d.dtab + (size_t)(t - 1) * L * C : nullptr;
          slot(k++) = q;
        }
      }
    }
    // net 1's tile groups start where net 0's end, so one phase pair spreads over all clusters
    for (size_t i = 1; i < ph.size(); i += 2) ph[i].goff = (s.ntiles * (ph[i - 1].NT / CS)) % ncl;
    SSB_CUDA(cudaMemcpyAsync(maps_dev, maps.data(), sizeof(CUtensorMap) * nmaps, cudaMemcpyHostToDevice, c.stream));
    SSB_CUDA(cudaMemcpyAsync(ph_dev, ph.data(), sizeof(SPhase) * nph, cudaMemcpyHostToDevice, c.stream));
    RUN(launch_sampler_tc(c, maps_dev, ph_dev, nph, s.tiles, s.tile_tight, s.ntiles, 2 * (2 * C / 64), ctr, CS));
  }
  c.release(mk);
  return 0;
}
bool f0_pair_persistent_ok(const Model& m, const SeqDev& s) {
  if (!m.persistent || !m.use_tc || s.ntiles > 48 || sampler_tc_max_ctas() <= 0) return false;
  for (int n = 0; n < 2; ++n) {
    const Denoiser& d = m.f0net[n];
    if (!denoiser_tc_ok(m, d) || !d.skip_tc.ok || !d.out_tc.ok || d.T <= 0) return false;
  }
  return m.f0net[0].C == m.f0net[1].C && m.f0net[0].L == m.f0net[1].L && m.f0net[0].T == m.f0net[1].T;
}

// a13+a14: GaussianMultinomialDiffusion.sample (gaussian_multinomial_diffusion.py:921-942)
int run_f0_diffusion(Ctx& c, const Model& m, int which, const SeqDev& s, const float* cond_g, const float* lo,
                     const float* hi, const float* gnoise, const float* unoise, uint64_t seed, float* z, int32_t* uv,
                     const Seq* host_seq) {
  const Denoiser& d = m.f0net[which];
  SSB_CHECK(d.T > 0, "f0 schedule not set: call ssb_model_set_schedule(which=1)");
  // EXPERIMENTAL utterance grouping, off unless SSB_F0_GROUP_FRAMES=<n> is set: see run_mel_diffusion.
  if (host_seq && !gnoise && !unoise && !c.dry) {
    const char* ge = getenv("SSB_F0_GROUP_FRAMES");
    const long gf = ge ? atol(ge) : 0;
    if (gf > 0 && host_seq->total > gf + gf / 2) {
      int b0 = 0, gi = 0;
      while (b0 < host_seq->B) {
        int b1 = b0;
        int64_t fr = 0;
        while (b1 < host_seq->B && (b1 == b0 || fr + host_seq->len[b1] <= gf)) fr += host_seq->len[b1++];
        std::vector<int32_t> offs((size_t)(b1 - b0) + 1, 0);
        for (int b = b0; b < b1; ++b) offs[(size_t)(b - b0) + 1] = offs[(size_t)(b - b0)] + host_seq->len[b];
        Seq q;
        q.build(offs.data(), b1 - b0);
        const size_t mkg = c.mark();
        SeqDev sg;
        RUN(upload_layout(c, q, 1, &sg));
        const int64_t ro = (int64_t)host_seq->rs[b0] - GUARD;  // all operands here are guard-banded [rows, ld] buffers
        RUN(run_f0_diffusion(c, m, which, sg, cond_g + ro * 256, lo + ro, hi + ro, nullptr, nullptr,
                             seed + 0x9E3779B97F4A7C15ull * (uint64_t)gi, z + ro, uv + ro, nullptr));
        c.release(mkg);
        b0 = b1;
        ++gi;
      }
      return 0;
    }
  }
  const size_t mk = c.mark();
  DenoiserBufs b;
  RUN(alloc_denoiser(c, d, s, denoiser_tc_ok(m, d), &b, m.cond_hoist));
  RUN(prepare_cond(c, d, s, cond_g, b));
  const int T = d.T;
  const size_t per = (size_t)s.total;
  const uint64_t sbase = 2000 + (uint64_t)which * 100000;
  RUN(f0_init(c, s, z, uv, gnoise, seed, sbase));
  for (int t = T - 1; t >= 0; --t) {
    const float* dt = d.dtab + (size_t)t * d.L * d.C;
    RUN(ddiff_input(c, s, z, uv, d.in_w, d.in_b, d.uv_emb, dt, b.tc ? nullptr : b.x, b.y, d.C, b.yh, b.yl));
    RUN(denoiser_stack(c, d, s, t, b));
    F0StepArgs a;
    a.z = z; a.uv = uv; a.out3 = b.head; a.ld3 = b.ld_head; a.lo = lo; a.hi = hi;
    a.gnoise = gnoise ? gnoise + per * (size_t)(T - t) : nullptr;
    a.unoise = unoise ? unoise + per * 2 * (size_t)(T - 1 - t) : nullptr;
    a.gtab = d.gtab + (size_t)t * 8; a.mtab = d.mtab + (size_t)t * 8; a.t = t; a.log_eps = m.log_eps;
    a.seed = seed; a.stream_id = sbase + 10 + 2 * (uint64_t)t;
    RUN(f0_p_sample(c, s, a));
  }
  c.release(mk);
  return 0;
}

// a13+a14 for BOTH F0 nets of a large batch in lock step (the agnostic and the specific sampler are independent,
// stylesinger.py:223-225): per reverse step the residual layers of the two nets run as the two lanes of the interleaved
// dual kernel, everything else (DDiffNet input, heads, Gaussian + multinomial update) per net as in run_f0_diffusion.
bool f0_dual_ok(const Model& m, const SeqDev& s) {
  if (!m.use_tc || !m.cond_hoist || !dual_enabled() || s.ntiles < 74) return false;
  for (int n = 0; n < 2; ++n) {
    const Denoiser& d = m.f0net[n];
    if (!denoiser_tc_ok(m, d) || !d.cond_all_tc.ok || d.T <= 0) return false;
    const int hb = d.layers[0].dil_tc.hb;
    if (hb != d.layers[0].outp_tc.hb || (hb != 128 && hb != 96)) return false;
  }
  return m.f0net[0].C == m.f0net[1].C && m.f0net[0].L == m.f0net[1].L && m.f0net[0].T == m.f0net[1].T &&
         m.f0net[0].layers[0].dil_tc.hb == m.f0net[1].layers[0].dil_tc.hb;
}
int run_f0_diffusion_dual(Ctx& c, const Model& m, const SeqDev& s, const float* cond0, const float* cond1, const float* lo,
                          const float* hi, const float* const gnoise[2], const float* const unoise[2], uint64_t seed,
                          float* const z[2], int32_t* const uv[2]) {
  const size_t mk = c.mark();
  DenoiserBufs b[2];
  const int T = m.f0net[0].T;
  const size_t per = (size_t)s.total;
  for (int n = 0; n < 2; ++n) {
    const Denoiser& d = m.f0net[n];
    RUN(alloc_denoiser(c, d, s, true, &b[n], true));
    RUN(prepare_cond(c, d, s, n == 0 ? cond0 : cond1, b[n]));
    RUN(f0_init(c, s, z[n], uv[n], gnoise[n], seed, 2000 + (uint64_t)n * 100000));
  }
  SSB_CHECK(c.dry || (b[0].condpre && b[1].condpre), "f0 dual sampler needs the hoisted conditioner");
  for (int t = T - 1; t >= 0; --t) {
    for (int n = 0; n < 2; ++n) {
      const Denoiser& d = m.f0net[n];
      RUN(ddiff_input(c, s, z[n], uv[n], d.in_w, d.in_b, d.uv_emb, d.dtab + (size_t)t * d.L * d.C, nullptr, b[n].y, d.C, b[n].yh, b[n].yl));
    }
    if (!c.dry) {
      const Lane A{&m.f0net[0], &b[0], s.rows, s.tiles, s.ntiles, t, 0}, B{&m.f0net[1], &b[1], s.rows, s.tiles, s.ntiles, t, 0};
      RUN(denoiser_layers_dual(c, A, B));
    }
    for (int n = 0; n < 2; ++n) {
      const Denoiser& d = m.f0net[n];
      RUN(denoiser_heads(c, d, s, b[n]));
      F0StepArgs a;
      a.z = z[n]; a.uv = uv[n]; a.out3 = b[n].head; a.ld3 = b[n].ld_head; a.lo = lo; a.hi = hi;
      a.gnoise = gnoise[n] ? gnoise[n] + per * (size_t)(T - t) : nullptr;
      a.unoise = unoise[n] ? unoise[n] + per * 2 * (size_t)(T - 1 - t) : nullptr;
      a.gtab = d.gtab + (size_t)t * 8; a.mtab = d.mtab + (size_t)t * 8; a.t = t; a.log_eps = m.log_eps;
      a.seed = seed; a.stream_id = 2000 + (uint64_t)n * 100000 + 10 + 2 * (uint64_t)t;
      RUN(f0_p_sample(c, s, a));
    }
  }
  c.release(mk);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// a20/a21: HifiGanGenerator.forward (hifigan_nsf.py:144-169)
// Stages whose channel counts are multiples of 64 run on the tcgen05 kernel: every conv input is carried as
// fp16 hi/lo planes of leaky_relu(x) written by the producing epilogue (the reference applies leaky_relu before
// every conv), residuals / MRF accumulators stay fp32.  The narrow stage (C = 32) runs its ResBlocks through a
// time-paired [rows/2, 64] view of the same memory with repacked weights (pack.cu, pack_conv_paired_tc).
int run_vocoder(Ctx& c, const Vocoder& v, const Seq& seq, const float* mel_tight, const float* f0_tight,
                const float* rand_ini, const float* src_noise, uint64_t seed, float* wav_tight) {
  const size_t mk0 = c.mark();
  int hop = 1;
  for (auto& st : v.stages) hop *= st.u;
  SeqDev s1, s256;
  RUN(upload_layout(c, seq, 1, &s1));
  RUN(upload_layout(c, seq, hop, &s256));
  float* mel = alloc_rows(c, s1, 80);
  float* f0g = alloc_rows(c, s1, 1);
  float* har = nullptr;
  WS_OK(c);
  RUN(pack_rows(c, s1, mel_tight, 80, mel, 80, 80));
  const bool nsf = v.nsf && f0_tight != nullptr;
  if (nsf) {
    RUN(pack_rows(c, s1, f0_tight, 1, f0g, 1, 1));
    har = alloc_rows(c, s256, 1);
    const size_t mk = c.mark();
    double* scratch = c.alloc<double>(nsf_scratch_doubles(s256));
    WS_OK(c);
    RUN(nsf_source(c, s1, s256, f0g, v.lin_w, v.lin_b, rand_ini, src_noise, har, scratch, seed, hop, (float)v.cfg.sample_rate));
    c.release(mk);
  }
  const bool tc = v.use_tc && tc_available();
  int C = v.cfg.initial_channel;
  float* x = alloc_rows(c, s1, C);
  __half *pin_h = nullptr, *pin_l = nullptr;  // planes of leaky_relu(stage input), when the next ups runs on tensor cores
  const bool up0_tc = tc && !v.stages.empty() && v.stages[0].up_tc.ok;
  if (up0_tc) {
    pin_h = alloc_half_rows(c, s1, C);
    pin_l = alloc_half_rows(c, s1, C);
  }
  WS_OK(c);
  {
    ConvGemm g = make_gemm(v.pre, s1, mel, 80);
    g.e.out = x; g.e.ldo = C;
    if (up0_tc) { g.e.out2_h = pin_h; g.e.out2_l = pin_l; g.e.ldh = C; g.e.plane_act = ACT_LRELU; g.e.plane_slope = 0.1f; }
    RUN(conv_gemm(c, g));
  }
  int rate = 1;
  SeqDev sin = s1;
  float* xin = x;
  for (size_t i = 0; i < v.stages.size(); ++i) {
    const VocStage& st = v.stages[i];
    const int Co = st.Cout;
    const int rate_out = rate * st.u;
    SeqDev so;
    RUN(upload_layout(c, seq, rate_out, &so));
    const bool up_tc = tc && st.up_tc.ok && pin_h != nullptr;
    const bool paired = tc && st.paired && (rate_out % 2 == 0);
    const bool res_tc = tc && (st.res_tc || paired);
    // paired stage: [rows, 32] is processed as [rows/2, 64] (same memory) with the time-paired weight packing
    SeqDev sw = so;
    const int Cw = paired ? 2 * Co : Co;
    if (paired) {
      RUN(upload_layout(c, seq, rate_out / 2, &sw));
      sw.rows = so.rows / 2;  // exactly the memory of the [so.rows, Co] buffers (TMA zero-fills beyond)
    }
    const bool next_up_tc = tc && i + 1 < v.stages.size() && v.stages[i + 1].up_tc.ok;
    float* xu = alloc_rows(c, so, Co);
    float* r = alloc_rows(c, so, Co);
    float* acc = alloc_rows(c, so, Co);
    float* xt = nullptr;
    __half *px_h = nullptr, *px_l = nullptr, *pt_h = nullptr, *pt_l = nullptr, *pr_h = nullptr, *pr_l = nullptr;
    __half *pa_h = nullptr, *pa_l = nullptr;
    if (res_tc) {
      px_h = alloc_half_rows(c, so, Co); px_l = alloc_half_rows(c, so, Co);
      pt_h = alloc_half_rows(c, so, Co); pt_l = alloc_half_rows(c, so, Co);
      pr_h = alloc_half_rows(c, so, Co); pr_l = alloc_half_rows(c, so, Co);
    } else {
      xt = alloc_rows(c, so, Co);
    }
    if (next_up_tc) { pa_h = alloc_half_rows(c, so, Co); pa_l = alloc_half_rows(c, so, Co); }
    WS_OK(c);
    // x = ups[i](leaky_relu(x, 0.1))
    if (up_tc) {
      GemmTC g;
      g.A_hi = pin_h; g.A_lo = pin_l; g.rows_total = sin.rows; g.w = &st.up_tc; g.tiles = sin.tiles; g.ntiles = sin.ntiles;
      g.e.mode = EPI_GENERIC; g.e.out = xu; g.e.ldo = st.u * Co;
      if (res_tc && !nsf) { g.e.oh = px_h; g.e.ol = px_l; g.e.ldh = st.u * Co; g.e.plane_act = ACT_LRELU; g.e.plane_slope = 0.1f; }
      RUN(conv_gemm_tc(c, g));
    } else {
      ConvGemm g = make_gemm(st.up, sin, xin, C);
      g.a_act = ACT_LRELU; g.a_slope = 0.1f;
      g.e.out = xu; g.e.ldo = st.u * Co;
      if (res_tc && !nsf) { g.e.out2_h = px_h; g.e.out2_l = px_l; g.e.ldh = st.u * Co; g.e.plane_act = ACT_LRELU; g.e.plane_slope = 0.1f; }
      RUN(conv_gemm(c, g));
    }
    if (nsf) RUN(noise_conv_add(c, so, s256, xu, Co, Co, har, st.nc_w, st.nc_b, st.nc_s, res_tc ? px_h : nullptr, px_l, 0.1f, st.nc_wt));
    for (int j = 0; j < v.nk; ++j) {  // MRF: mean of the resblocks
      const float* rin = xu;
      const __half *rin_h = px_h, *rin_l = px_l;
      for (int mI = 0; mI < 3; ++mI) {
        const bool last = (mI == 2);
        const bool lastj = (j == v.nk - 1);
        if (res_tc) {
          {
            GemmTC g;  // xt = c1(leaky_relu(r)) ; only leaky_relu(xt) is ever consumed -> planes only
            g.A_hi = rin_h; g.A_lo = rin_l; g.rows_total = sw.rows; g.w = &st.rb[j].c1_tc[mI]; g.tiles = sw.tiles; g.ntiles = sw.ntiles;
            g.e.mode = EPI_GENERIC; g.e.oh = pt_h; g.e.ol = pt_l; g.e.ldh = Cw; g.e.plane_act = ACT_LRELU; g.e.plane_slope = 0.1f;
            RUN(conv_gemm_tc(c, g));
          }
          GemmTC g;  // r = c2(leaky_relu(xt)) + r
          g.A_hi = pt_h; g.A_lo = pt_l; g.rows_total = sw.rows; g.w = &st.rb[j].c2_tc[mI]; g.tiles = sw.tiles; g.ntiles = sw.ntiles;
          g.e.mode = EPI_GENERIC; g.e.res = rin; g.e.ld_res = Cw;
          if (!last) {
            g.e.out = r; g.e.ldo = Cw; g.e.oh = pr_h; g.e.ol = pr_l; g.e.ldh = Cw; g.e.plane_act = ACT_LRELU; g.e.plane_slope = 0.1f;
          } else {
            g.e.out = acc; g.e.ldo = Cw; g.e.accum = (j > 0); g.e.gamma = lastj ? 1.0f / (float)v.nk : 1.0f;
            if (lastj && next_up_tc) { g.e.oh = pa_h; g.e.ol = pa_l; g.e.ldh = Cw; g.e.plane_act = ACT_LRELU; g.e.plane_slope = 0.1f; }
          }
          RUN(conv_gemm_tc(c, g));
          rin = r; rin_h = pr_h; rin_l = pr_l;
          continue;
        }
        {
          ConvGemm g = make_gemm(st.rb[j].c1[mI], so, rin, Co);
          g.a_act = ACT_LRELU; g.a_slope = 0.1f;
          g.e.out = xt; g.e.ldo = Co;
          RUN(conv_gemm(c, g));
        }
        ConvGemm g = make_gemm(st.rb[j].c2[mI], so, xt, Co);
        g.a_act = ACT_LRELU; g.a_slope = 0.1f;
        g.e.res = rin; g.e.ld_res = Co;
        if (!last) {
          g.e.out = r; g.e.ldo = Co;
        } else {
          g.e.out = acc; g.e.ldo = Co;
          g.e.accum = (j > 0);
          g.e.gamma = lastj ? 1.0f / (float)v.nk : 1.0f;
          if (lastj && next_up_tc) { g.e.out2_h = pa_h; g.e.out2_l = pa_l; g.e.ldh = Co; g.e.plane_act = ACT_LRELU; g.e.plane_slope = 0.1f; }
        }
        RUN(conv_gemm(c, g));
        rin = r;
      }
    }
    xin = acc; sin = so; C = Co; rate = rate_out;
    pin_h = pa_h; pin_l = pa_l;
  }
  if (v.post.N == 1 && C % 4 == 0 && (size_t)((256 + v.post.taps - 1) * (C + 1) + v.post.taps * C) * 4 <= 48 * 1024) {
    // N = 1: dedicated windowed reduction (leaky_relu 0.01 -> conv_post -> tanh in one pass over x)
    RUN(conv_post_tanh(c, sin, xin, C, C, v.post.taps, v.post.center, v.post.W, v.post.Npad, v.post.bias, 0.01f, wav_tight));
  } else {
    float* y = alloc_rows(c, sin, 4);
    WS_OK(c);
    ConvGemm g = make_gemm(v.post, sin, xin, C);
    g.a_act = ACT_LRELU; g.a_slope = 0.01f;  // F.leaky_relu default slope (hifigan_nsf.py:165)
    g.e.out = y; g.e.ldo = 4;
    RUN(conv_gemm(c, g));
    RUN(tanh_out(c, sin, y, 4, wav_tight));
  }
  c.release(mk0);
  return 0;
}

}  // namespace ssb
