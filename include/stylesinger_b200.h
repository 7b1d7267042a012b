/* stylesinger_b200 — C ABI of the B200-native StyleSinger hot path (libstylesinger_b200.so).
 *
 * The reference (AaronZ345/StyleSinger) is pure Python/PyTorch and has NO FFI of its own
 * (SURVEY.md §8b); its extension points are Python classes and registries.  This header is the
 * boundary a binding for that path attaches to: each entry point replaces one reference interface,
 * cited as file:line of /root/reference.  INTEGRATION.md shows the ctypes stubs and the
 * reference-side registrations (DIFF_DECODERS / FS_ENCODERS / FS_DECODERS / register_vocoder).
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.  `stream` is a cudaStream_t passed as void*.
 *  - all DEVICE tensors are fp32 (or int32) row-major and "tight packed": a batch of B utterances
 *    of lengths L_b is one [sum L_b, C] matrix; `*_offsets` are HOST int32 arrays of B+1 prefix
 *    sums.  Every utterance is processed with true-length (the reference's B=1) semantics —
 *    results do not depend on batch composition.
 *  - the caller owns every device buffer including the workspace (`*_workspace_bytes`); entry
 *    points enqueue on `stream`, never allocate device memory and never synchronise — except
 *    ssb_model_create / ssb_vocoder_create / ssb_model_set_schedule, which own the packed weights.
 *  - return 0 on success, negative on error with a message in ssb_last_error() (thread-local).
 *  - noise: NULL noise pointers select the in-kernel counter-based generator (Philox, `seed`);
 *    non-NULL pointers inject the noise explicitly (parity mode; SURVEY.md A.10 draw order).
 */
#ifndef STYLESINGER_B200_H
#define STYLESINGER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ssb_model ssb_model_t;     /* packed StyleSinger acoustic model (immutable after create) */
typedef struct ssb_vocoder ssb_vocoder_t; /* packed HiFi-GAN(-NSF) generator */
typedef struct ssb_melspec ssb_melspec_t; /* STFT + mel filterbank of the reference-audio front-end */
typedef struct ssb_lstm_encoder ssb_lstm_encoder_t; /* LSTM utterance encoder of the reference-audio front-end (emo_embed) */

/* One named fp32 HOST tensor of a reference state_dict (names exactly as in the reference's
 * checkpoints: utils/commons/ckpt_utils.py:26-67 loads state_dict['model']). */
typedef struct {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} ssb_tensor_desc;

/* egs/stylesinger.yaml keys the kernels are specialised on (reference egs/stylesinger.yaml:10-141). */
typedef struct {
  int32_t hidden_size;     /* 256 */
  int32_t enc_layers, dec_layers;           /* 4, 4 */
  int32_t enc_ffn_kernel, dec_ffn_kernel;   /* 9, 9 */
  int32_t dur_layers, dur_kernel;           /* 2, 3 */
  int32_t n_tokens;                         /* len(phone dictionary) */
  int32_t n_rq, rq_depth;                   /* 128, 4 */
  int32_t mel_channels, mel_layers, mel_cycle; /* residual_channels 256, residual_layers 20, dilation_cycle_length 4 */
  int32_t f0_channels, f0_layers, f0_cycle;    /* 192, 10, 4 */
  int32_t mel_bins;                         /* 80 */
} ssb_hparams;

/* HiFi-GAN generator config (checkpoints/hifigan/config.yaml keys read at
 * tasks/tts/vocoder_infer/hifigan_nsf.py:48-60, modules/hifigan/hifigan_nsf.py:104-142). */
typedef struct {
  int32_t n_up;
  int32_t up_rates[8];
  int32_t up_kernels[8];
  int32_t initial_channel;
  int32_t n_res;
  int32_t res_kernels[4];
  int32_t res_dilations[4][3];
  int32_t use_pitch_embed; /* NSF harmonic source */
  int32_t sample_rate;
} ssb_vocoder_config;

int ssb_version(void);
const char* ssb_last_error(void);

/* Replaces StyleSinger.__init__ + load_ckpt (modules/StyleSinger/stylesinger.py:46-117,
 * utils/commons/ckpt_utils.py:26-67).  Extra host-computed constant expected in `tensors`:
 *   "__pos_table" [rows,256]: SinusoidalPositionalEmbedding.get_embedding(rows,256,0)
 *   (modules/commons/common_layers.py:111-127), rows >= max sequence length + 2. */
int ssb_model_create(ssb_model_t** out, const ssb_tensor_desc* tensors, int32_t n, const ssb_hparams* hp);
void ssb_model_free(ssb_model_t* m);

/* Diffusion schedules (GaussianDiffusion.__init__ modules/diff/shallow_diffusion_tts.py:68-122;
 * GaussianMultinomialDiffusion.__init__ modules/diff/gaussian_multinomial_diffusion.py:208-284).
 * which: 0 = mel denoiser, 1 = both F0 denoisers.  Host arrays:
 *   step_emb [T, C]   SinusoidalPosEmb(t) (modules/diff/net.py:31-44), C = residual channels
 *   gauss_tab [T, 8]  {sqrt_recip_ac, sqrt_recipm1_ac, post_coef1, post_coef2, sigma, sqrt_ac, sqrt_1m_ac, 0}
 *   multi_tab [T, 8]  {log_alpha_t, log_1m_alpha_t, log_cumprod_alpha_{t-1}, log_1m_cumprod_alpha_{t-1}, ...} (which==1)
 * Builds the per-layer step-bias table [T, L, C] on the device.  Synchronises `stream`. */
int ssb_model_set_schedule(ssb_model_t* m, int32_t which, int32_t T, const float* step_emb, const float* gauss_tab,
                           const float* multi_tab, void* stream);

/* Inputs of StyleSinger.forward(..., infer=True) (modules/StyleSinger/stylesinger.py:119-187). */
typedef struct {
  int32_t B;
  const int32_t* ph_offsets;    /* host [B+1] */
  const int32_t* frame_offsets; /* host [B+1]; required by ssb_acoustic_forward */
  const int32_t* ref_offsets;   /* host [B+1] */
  const int32_t* txt_tokens;    /* dev [sumP] */
  const int32_t* note;          /* dev [sumP] */
  const int32_t* note_type;     /* dev [sumP] */
  const float* note_dur;        /* dev [sumP] */
  const float* spk_embed;       /* dev [B,256] */
  const float* emo_embed;       /* dev [B,256] */
  const float* ref_mels;        /* dev [sumR,80] */
  const float* ref_f0;          /* dev [sumR] */
  const int32_t* mel2ph;        /* dev [sumF] 1-based phone index per frame, or NULL -> use `dur` */
  const int32_t* dur;           /* dev [sumP] frames per phone (from ssb_predict_durations), or NULL */
  const float* f0;              /* dev [sumF] optional teacher-forced log2-Hz f0 (forward kwarg f0) */
  const float* uv;              /* dev [sumF] optional teacher-forced uv */
  const float* f0_gauss_noise[2]; /* dev [(T_f0+1), sumF]: z init then one per step t=T-1..0; NULL -> Philox */
  const float* f0_unif_noise[2];  /* dev [T_f0, sumF, 2] */
  const float* mel_noise;         /* dev [(T+1), sumF, 80]: q_sample then one per step */
  uint64_t seed;
  int32_t skip_mel_diffusion;     /* 1: stop after the coarse mel / diff_cond */
  int32_t pndm_speedup;           /* 0: DDPM ancestral sampling, T steps (the StyleSinger default, DiffusionDecoder.forward);
                                   * k > 0: PLMS with iteration interval k (hparams['pndm_speedup'],
                                   * modules/diff/shallow_diffusion_tts.py:164-197,254-260): T / k (+1) denoiser evaluations */
} ssb_acoustic_inputs;

/* Outputs (all optional except mel_out/f0_denorm when diffusion runs); device, tight packed. */
typedef struct {
  float* mel_out;       /* [sumF,80]   ret['mel_out'] */
  float* f0_denorm;     /* [sumF]      ret['f0_denorm'] (Hz) */
  float* encoder_out;   /* [sumP,256]  encoder(txt)+note_encoder */
  float* style;         /* [sumF,256]  ret['style'] */
  int32_t* rq_codes;    /* [sumR,4]    RVQ indices */
  float* pitch_pred;    /* [sumF,2]    ret['pitch_pred'] */
  float* decoder_inp;   /* [sumF,256]  ret['decoder_inp'] */
  float* coarse_mel;    /* [sumF,80]   FFT-decoder mel before diffusion */
  float* diff_cond;     /* [sumF,256]  ln_proj(cat[...]) */
  int32_t* mel2ph;      /* [sumF] */
  float* spk_proj;      /* [B,256] ret['spk_embed'] */
  float* emo_proj;      /* [B,256] ret['emo_embed'] */
} ssb_acoustic_outputs;

/* FastSpeech2.add_dur -> DurationPredictor.inference (modules/fastspeech/fs2.py:151-174,
 * modules/fastspeech/tts_modules.py:105-130): dur[p] = clamp(round(exp(x)-1), 0).
 * The host reads `dur_out` back to size the frame axis (the reference syncs here too). */
size_t ssb_durations_workspace_bytes(const ssb_model_t* m, const ssb_acoustic_inputs* in);
int ssb_predict_durations(const ssb_model_t* m, const ssb_acoustic_inputs* in, int32_t* dur_out, float* logdur_out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* StyleSinger.forward(infer=True, global_steps > diff_start): rows a1-a19 of SURVEY.md §8. */
size_t ssb_acoustic_workspace_bytes(const ssb_model_t* m, const ssb_acoustic_inputs* in);
int ssb_acoustic_forward(const ssb_model_t* m, const ssb_acoustic_inputs* in, const ssb_acoustic_outputs* out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* DiffusionDecoder.forward(infer=True) alone (modules/diff/shallow_diffusion_tts.py:284-307):
 * cond [sumF,256], coarse [sumF,80] -> mel [sumF,80]. */
size_t ssb_mel_diffusion_workspace_bytes(const ssb_model_t* m, const int32_t* frame_offsets, int32_t B);
int ssb_mel_diffusion_sample(const ssb_model_t* m, const float* cond, const float* coarse_mel,
                             const int32_t* frame_offsets, int32_t B, const float* noise, uint64_t seed,
                             float* mel_out, void* workspace, size_t workspace_bytes, void* stream);

/* PLMS / PNDM sampler over the same DiffNet: GaussianDiffusion.p_sample_plms driven by the
 * `pndm_speedup` loop of GaussianDiffusion.forward (modules/diff/shallow_diffusion_tts.py:164-197,254-260).
 * interval = hparams['pndm_speedup']; q_noise [sumF,80] (tight) is the single q_sample draw, NULL = in-kernel Philox. */
size_t ssb_mel_diffusion_plms_workspace_bytes(const ssb_model_t* m, const int32_t* frame_offsets, int32_t B);
int ssb_mel_diffusion_sample_plms(const ssb_model_t* m, const float* cond, const float* coarse_mel,
                                  const int32_t* frame_offsets, int32_t B, const float* q_noise, uint64_t seed,
                                  int32_t interval, float* mel_out, void* workspace, size_t workspace_bytes, void* stream);

/* One denoiser evaluation, DiffNet.forward / DDiffNet.forward (modules/diff/net.py:107-130,242-266).
 * which: 0 mel (x [sumF,80] -> eps [sumF,80]); 1 / 2 F0 agnostic / specific (x = f0 [sumF], uv int32 [sumF]
 * -> out [sumF,3]). */
int ssb_denoiser_eval(const ssb_model_t* m, int32_t which, const float* x, const int32_t* uv, int32_t t,
                      const float* cond, const int32_t* frame_offsets, int32_t B, float* out, void* workspace,
                      size_t workspace_bytes, void* stream);

/* GaussianMultinomialDiffusion.sample (modules/diff/gaussian_multinomial_diffusion.py:921-942).
 * which: 0 agnostic net (gm_diffnet), 1 specific (gm_diffnet_inpainte). cond [sumF,256], clip lo/hi [sumF]
 * -> f0_norm [sumF] (normalised), uv int32 [sumF]. */
int ssb_f0_diffusion_sample(const ssb_model_t* m, int32_t which, const float* cond, const float* clip_lo,
                            const float* clip_hi, const int32_t* frame_offsets, int32_t B, const float* gauss_noise,
                            const float* unif_noise, uint64_t seed, float* f0_norm_out, int32_t* uv_out,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Per-registry drop-ins.  The reference dispatches through FS_ENCODERS / FS_DECODERS
 * (modules/fastspeech/fs2.py:9-18,30-31) and calls StyleSinger.get_style (modules/StyleSinger/stylesinger.py:189-214):
 *  ssb_fft_encoder  = FastspeechEncoder.forward(txt_tokens) (tts_modules.py:326-346): tokens int32 [sumP] -> [sumP,256]
 *                     (embedding * sqrt(H) + sinusoidal positions, FFT blocks, final LayerNorm, padding rows zero);
 *  ssb_fft_decoder  = FastspeechDecoder.forward(x) (tts_modules.py:349-355, FFTBlocks.forward :281-306): x [sumF,256]
 *                     -> [sumF,256], padding mask = rows of x that are all zero;
 *  ssb_get_style    = get_style(decoder_inp, ref_mels, ...) with ret['ref_f0'] (LocalStyleAdaptor + RVQ + l1 +
 *                     ProsodyAligner, lse.py:103-129,59-81): -> style [sumF,256], codes int32 [sumR,depth] (optional).
 * which (workspace query): 0 encoder (offsets = ph_offsets), 1 decoder (offsets = frame_offsets). */
size_t ssb_fft_workspace_bytes(const ssb_model_t* m, int32_t which, const int32_t* offsets, int32_t B);
int ssb_fft_encoder(const ssb_model_t* m, const int32_t* txt_tokens, const int32_t* ph_offsets, int32_t B, float* out,
                    void* workspace, size_t workspace_bytes, void* stream);
int ssb_fft_decoder(const ssb_model_t* m, const float* x, const int32_t* frame_offsets, int32_t B, float* out,
                    void* workspace, size_t workspace_bytes, void* stream);
size_t ssb_get_style_workspace_bytes(const ssb_model_t* m, const int32_t* frame_offsets, const int32_t* ref_offsets, int32_t B);
int ssb_get_style(const ssb_model_t* m, const float* decoder_inp, const int32_t* frame_offsets, const float* ref_mels,
                  const float* ref_f0, const int32_t* ref_offsets, int32_t B, float* style_out, int32_t* codes_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* RQBottleneck.forward / VQEmbedding.find_nearest_embedding (modules/StyleSinger/RQ.py:262-270,29-55). */
int ssb_rvq_lookup(const ssb_model_t* m, const float* x /*[sumR,256]*/, const int32_t* ref_offsets, int32_t B,
                   float* quant_out /*[sumR,256]*/, int32_t* codes_out /*[sumR,depth]*/, void* workspace,
                   size_t workspace_bytes, void* stream);

/* HifiGanGenerator (modules/hifigan/hifigan_nsf.py:104-178) with weight norm folded at pack time. */
int ssb_vocoder_create(ssb_vocoder_t** out, const ssb_tensor_desc* tensors, int32_t n, const ssb_vocoder_config* cfg);
void ssb_vocoder_free(ssb_vocoder_t* v);

/* HifiGAN.spec2wav / HifiGanGenerator.forward (tasks/tts/vocoder_infer/hifigan_nsf.py:62-75,
 * modules/hifigan/hifigan_nsf.py:144-169).  mel [sumF,80] (already masked/clipped as
 * inference/StyleSinger.py:56-58 does), f0 [sumF] Hz or NULL.  rand_ini [B,9] / src_noise [sumF*hop, 9]:
 * SineGen's torch.rand / torch.randn draws (modules/parallel_wavegan/models/source.py:357,436) or NULL.
 * wav_out [sumF*hop]. */
size_t ssb_vocoder_workspace_bytes(const ssb_vocoder_t* v, const int32_t* frame_offsets, int32_t B);
int ssb_hifigan_generate(const ssb_vocoder_t* v, const float* mel, const float* f0, const int32_t* frame_offsets,
                         int32_t B, const float* rand_ini, const float* src_noise, uint64_t seed, float* wav_out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Select the GEMM path of the denoiser layers: 1 = tcgen05 tensor cores on fp16 hi/lo split operands
 * (3 MMAs per product, fp32 accumulate; default when available), 0 = fp32 FFMA.  Returns the mode in effect. */
int ssb_model_set_tensor_cores(ssb_model_t* m, int32_t enable);

/* Same switch for the vocoder's wide stages (C % 64 == 0). */
int ssb_vocoder_set_tensor_cores(ssb_vocoder_t* v, int32_t enable);

/* 1 (default): small batches run the whole T-step mel sampler in ONE persistent cooperative kernel launch
 * (csrc/sampler_tc.cu); 0: one launch per GEMM (BASELINE.json configs[4] compares the two). */
int ssb_model_set_persistent(ssb_model_t* m, int32_t enable);
/* 0 (default): batches of more than 48 row tiles (~6 k frames) take the one-launch-per-GEMM path.  1: such batches are
 * split into groups of consecutive utterances of <= 48 row tiles and every group runs the T-step mel sampler
 * (shallow_diffusion_tts.py:303-304) as ONE persistent launch (production RNG mode only; each group draws from its own
 * Philox stream).  This is the "persistent-kernel" arm of BASELINE.json configs[4] at batch 64. */
int ssb_model_set_persistent_groups(ssb_model_t* m, int32_t enable);
/* 1 (default): the per-launch tcgen05 samplers compute the step-invariant conditioner_projection of all residual layers
 * (modules/diff/net.py:59,71 - Conv1d(256, 2C, 1) applied to the same cond in every one of the T steps) ONCE per sampler call
 * and add it in the gate epilogue; 0: contract it inside every layer GEMM of every step (round-1 behaviour). */
int ssb_model_set_cond_hoist(ssb_model_t* m, int32_t enable);
/* Decoder FFT blocks (modules/commons/transformer.py TransformerFFNLayer, conv k=9 -> gelu -> linear): run the FFN GEMMs
 * on the tcgen05 kernel for batches of >= 1024 frames (default on; 0 keeps them on the fp32 FFMA kernel). Returns the
 * new setting. */
int ssb_model_set_fft_tensor_cores(ssb_model_t* m, int32_t enable);

/* Unit-test granularity: ssb_op_conv1d through the tcgen05 path (Cin % 64 == 0, N % 128 == 0, no activation). */
int ssb_op_conv1d_tc(const float* x, const int32_t* offsets, int32_t B, int32_t Cin, const float* w_host,
                     const float* b_host, int32_t N, int32_t k, int32_t dilation, float* out, void* stream);

/* StyleSingerInfer.forward_model glue between model and vocoder (inference/StyleSinger.py:56-58):
 * clips mel [n_frames,80] in place to [vmin, vmax] and counts the frames with sum|mel| > 0 into
 * *nonzero_frames (device int32; the reference drops all-zero frames, which only padding can produce). */
int ssb_mel_postprocess(float* mel, int64_t n_frames, float vmin, float vmax, int32_t* nonzero_frames, void* stream);

/* ---- f3: mel-spectrogram of the reference audio -------------------------------------------------
 * Replaces utils/audios/__init__.py:36-84 librosa_wav2spec as called by inference/StyleSinger.py:79-92 (process_audio):
 * librosa.stft(center=True, pad_mode="constant", periodic Hann window) -> |.| -> librosa.filters.mel (Slaney scale and
 * normalisation, built inside create) -> log10(max(eps, .)).  fmin / fmax < 0 mean 0 / sample_rate / 2 like the reference.
 * Constraints of the implicit-GEMM formulation: fft_size even, fft_size <= 32 hop_size (16 with reflect centring), hop_size a multiple of 16, n_mels of 4
 * (egs/stylesinger.yaml: 48 kHz, fft 1024, hop 256, win 1024, 80 mels, 20..24000 Hz).  An utterance of n samples yields
 * ssb_melspec_num_frames = 1 + n / hop_size frames.  wav: device fp32, utterances concatenated, sample_offsets: host [B+1];
 * mel_out: device fp32 [sum frames, n_mels].  loud_norm / trim_long_sil of the reference are not implemented (both false in
 * the reference's configuration); the speaker / emotion encoders and the Praat pitch tracker are outside this library. */
int ssb_melspec_create(ssb_melspec_t** out, int32_t sample_rate, int32_t fft_size, int32_t hop_size, int32_t win_length,
                       int32_t n_mels, float fmin, float fmax, float eps);
/* Same front-end with the defaults of librosa.feature.melspectrogram selectable, which is what the emotion encoder's features
 * are (data_gen/tts/emotion/audio.py:43-55: sr 16000, n_fft 400, hop 160, 40 mels, fmin 0, fmax sr / 2): pad_reflect != 0 =
 * np.pad "reflect" centring instead of zeros (needs more than n_fft / 2 samples per utterance, like numpy), power != 0 = |X|^2
 * instead of |X|, take_log == 0 = no log10 / eps.  n_fft need not be a multiple of hop_size here (the frame is embedded in
 * the next multiple of 2 hop_size rows with zero weights).  ssb_melspec_create = (..., 0, 0, 1). */
int ssb_melspec_create_ex(ssb_melspec_t** out, int32_t sample_rate, int32_t fft_size, int32_t hop_size, int32_t win_length,
                          int32_t n_mels, float fmin, float fmax, float eps, int32_t pad_reflect, int32_t power, int32_t take_log);
void ssb_melspec_free(ssb_melspec_t* m);
int32_t ssb_melspec_num_frames(const ssb_melspec_t* m, int64_t n_samples);
size_t ssb_melspec_workspace_bytes(const ssb_melspec_t* m, const int32_t* sample_offsets, int32_t B);
int ssb_melspec_forward(const ssb_melspec_t* m, const float* wav, const int32_t* sample_offsets, int32_t B, float* mel_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- f3: LSTM utterance encoder of the reference audio -------------------------------------------
 * Replaces data_gen/tts/emotion/model.py:10-77 (EmotionEncoder: batch-first torch.nn.LSTM 40 -> 256 x 3 layers from zero
 * state + linear 256 -> 256) and the aggregation of data_gen/tts/emotion/inference.py:43-55 (embed_frames_batch) and
 * :150-151 (embed_utterance: mean of the partial embeddings, L2-normalised), which produce the `emo_embed` input of
 * inference/StyleSinger.py:106.  create: HOST fp32 pointers in torch state_dict layout, one per layer -
 * weight_ih[l] [4H, input_size or H], weight_hh[l] [4H, H], bias_ih[l] / bias_hh[l] [4H], gate rows i | f | g | o;
 * linear_weight [embed_size, H] / linear_bias [embed_size] may both be NULL (no `forward` head).  hidden_size must be 256.
 * forward: frames device fp32 [n_partials, n_frames, input_size] (every partial the same length, as the reference batches
 * them).  Outputs (device fp32, each may be NULL, at least one not): hidden_out [n_partials, H] = EmotionEncoder.inference
 * (hidden[-1]); embeds_out [n_partials, embed_size] = EmotionEncoder.forward; utt_embed_out [n_utterances, H] = normalised
 * mean of hidden over partials utt_offsets[u] .. utt_offsets[u+1] (host int32 [n_utterances + 1], 0 .. n_partials, every
 * utterance non-empty; only read when utt_embed_out is given). */
int ssb_lstm_encoder_create(ssb_lstm_encoder_t** out, int32_t input_size, int32_t hidden_size, int32_t num_layers,
                            const float* const* weight_ih, const float* const* weight_hh, const float* const* bias_ih,
                            const float* const* bias_hh, int32_t embed_size, const float* linear_weight, const float* linear_bias);
void ssb_lstm_encoder_free(ssb_lstm_encoder_t* m);
size_t ssb_lstm_encoder_workspace_bytes(const ssb_lstm_encoder_t* m, int32_t n_partials, int32_t n_frames, int32_t n_utterances);
int ssb_lstm_encoder_forward(const ssb_lstm_encoder_t* m, const float* frames, int32_t n_partials, int32_t n_frames,
                             const int32_t* utt_offsets, int32_t n_utterances, float* hidden_out, float* embeds_out,
                             float* utt_embed_out, void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernels this library has launched in this process so far (bench.py reports the delta). */
int64_t ssb_launch_count(void);

/* Diagnostics of the tensor-core GEMM dispatcher (csrc/conv_gemm_tc.cu): launches of one kernel variant, named
 * "tc<BN,MODE>" (single CTA) or "tc2<HB,MODE>" (CTA pair), MODE in GATE / RES_SKIP / GENERIC, e.g. "tc2<64,GENERIC>";
 * the ';'-separated list of variants launched so far (returns their number); and the counters of the activation
 * TMA-descriptor cache (descriptors encoded / served from the cache).  Tests use these to assert WHICH kernel a problem
 * size took; nothing in the reference corresponds to them. */
int64_t ssb_variant_launch_count(const char* variant);
int32_t ssb_variant_names(char* buf, int32_t cap);
void ssb_tensor_map_cache_stats(int64_t* encodes, int64_t* hits);

/* Process-wide switch (default 0) of the interleaved residual-layer schedule of the large-batch samplers: the gate conv of
 * one group of utterances (or of one F0 net) and the 1x1 residual/skip conv of the other group share one launch and
 * alternate tile by tile inside every SM (kernel variant "tc2d<HB,GATE+RES_SKIP>"; reference loop net.py:66-78 inside
 * shallow_diffusion_tts.py:303-304 / gaussian_multinomial_diffusion.py:928-939).  Results agree with one launch per GEMM to
 * fp32 rounding (tests/test_gpu_scale.py); measured 3 % slower than it on B200, hence opt-in.
 * Returns the new state. */
int32_t ssb_set_interleaved_layers(int32_t enable);

/* Process-wide switch of the tcgen05 / TMA attention kernel (csrc/attention_tc.cu) for the long-batch paths of the FFT blocks
 * (common_layers.py:277-286) and of the style aligner's cross-attention (lse.py:41); short batches always use the fp32
 * kernel.  Returns the new state. */
int32_t ssb_set_attention_tensor_cores(int32_t enable);

/* Unit-test granularity: one Conv1d over ragged rows with torch-layout HOST weights [N,Cin,k]
 * (packs on the fly with cudaMalloc; not for production use).  act: 0 none 1 relu 2 gelu 3 leaky(0.1) 4 tanh. */
int ssb_op_conv1d(const float* x, const int32_t* offsets, int32_t B, int32_t Cin, const float* w_host,
                  const float* b_host, int32_t N, int32_t k, int32_t dilation, int32_t act, float* out, void* stream);
/* Unit-test granularity: multi-head attention, 2 heads x 128; q [sumL,256], k/v [sumS,256]. */
int ssb_op_attention(const float* q, const float* k, const float* v, const int32_t* q_offsets,
                     const int32_t* k_offsets, int32_t B, float scale, float* out, void* stream);
/* Same contract on the tcgen05 / TMA attention kernel (csrc/attention_tc.cu; 3-pass fp16 hi/lo split MMAs for QK^T and PV,
 * TMA-staged K / V^T tiles), the path long batches take inside the FFT blocks (common_layers.py:277-286) and the style
 * aligner (lse.py:41). */
int ssb_op_attention_tc(const float* q, const float* k, const float* v, const int32_t* q_offsets,
                        const int32_t* k_offsets, int32_t B, float scale, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STYLESINGER_B200_H */
