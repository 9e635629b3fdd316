/*
 * ttsb200 — C ABI of the B200-native synthesis engine for illufly-tts.
 *
 * This is the drop-in boundary for ONE path of the reference: the engine forward pass behind
 * `BatchKModel` (/root/reference/src/illufly_tts/core/kmodel.py:5-150), i.e.
 *     phoneme ids + voice style vector  ->  24 kHz waveform.
 * The reference reaches that path through the Python class `BatchKModel(kokoro.model.KModel)`;
 * the arithmetic lives in the un-vendored PyPI package kokoro==0.9.4 (poetry.lock:939-955).
 * Every entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary;
 *   - every function returns 0 on success or a negative TTSB_E_* code and never throws;
 *     `ttsb_last_error()` gives a human-readable message for the calling thread's last failure;
 *   - "host" pointers are ordinary CPU memory, "dev" pointers are CUDA device memory on the
 *     engine's device; the engine never keeps a caller pointer after the call returns;
 *   - ragged batches are described CSR-style: `token_offsets[B+1]`, `frame_offsets[B+1]`;
 *   - there is NO CPU fallback: without a CUDA device `ttsb_engine_create` fails.
 */
#ifndef TTSB200_H_
#define TTSB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTSB_ABI_VERSION 1

#if defined(__GNUC__)
#define TTSB_API __attribute__((visibility("default")))
#else
#define TTSB_API
#endif

enum {
  TTSB_OK = 0,
  TTSB_E_INVALID = -1,   /* bad argument / unknown tensor / shape mismatch            */
  TTSB_E_CUDA = -2,      /* a CUDA runtime call or kernel launch failed                */
  TTSB_E_STATE = -3,     /* call order violated (e.g. finish() without begin())        */
  TTSB_E_NOMEM = -4,     /* workspace could not be allocated                           */
  TTSB_E_TOO_LONG = -5,  /* an utterance exceeds 512 tokens (ALBERT positions; kokoro asserts the same), or the batch
                            needs more frames than a call can address */
  TTSB_E_RANGE = -6      /* an operand of the split-fp16 (fp32-class) tensor-core path left fp16's range: the result would
                            not be fp32-class; rerun with TTSB_PRECISION_FP32_FFMA */
};

/* arithmetic of the convolutions / projections.  The text side, which decides the integer durations, has fp32-class
 * accuracy in every mode (tcgen05 with split-fp16 operands: x = hi + lo, three MMAs per K step, fp32 accumulation; or fp32
 * FFMA in TTSB_PRECISION_FP32_FFMA); pred_dur is bit-exact against the fp32 reference in all four.  The modes differ in the
 * frame side (decoder + vocoder): */
enum {
  TTSB_PRECISION_FP32 = 0,   /* the strict tier: every convolution on tcgen05 with split-fp16 operands (fp32-class
                                products): waveform within 1e-3 max-abs of the fp32 reference                      */
  TTSB_PRECISION_BF16 = 1,   /* tcgen05 (kind::f16) bf16 operands, fp32 accumulate/statistics                    */
  TTSB_PRECISION_FP16 = 2,   /* tcgen05 (kind::f16) fp16 operands (saturating), fp32 accumulate/statistics: same
                                instruction and rate as bf16, 3 more mantissa bits (~ +18 dB SNR)                */
  TTSB_PRECISION_FP32_FFMA = 3 /* CUDA-core fp32 FFMA everywhere: the arithmetic reference of the strict tier (tests), and
                                the mode for checkpoints whose activations leave the split path's range (TTSB_E_RANGE) */
};

enum {
  TTSB_MEM_HOST = 0,
  TTSB_MEM_DEVICE = 1
};

typedef struct ttsb_engine ttsb_engine;

TTSB_API int ttsb_abi_version(void);
/* message of the last failure on this thread ("" if none) */
TTSB_API const char* ttsb_last_error(void);

/* ---- lifetime ------------------------------------------------------------------------------------
 * replaces: BatchKModel(repo_id).to(device).eval()           (core/pipeline.py:70)
 *           kokoro KModel.__init__ checkpoint load            (SURVEY.md §5 "Checkpoint / resume")   */
TTSB_API int ttsb_engine_create(int cuda_device, ttsb_engine** out);
TTSB_API int ttsb_engine_destroy(ttsb_engine* e);

/* Hand the engine one tensor of the checkpoint, under its checkpoint key
 * "<module>.<state-dict key without the 'module.' prefix>", e.g.
 *   "decoder.generator.ups.0.weight_g", "predictor.lstm.weight_hh_l0_reverse",
 *   "bert.encoder.albert_layer_groups.0.albert_layers.0.attention.query.weight".
 * `data` is host fp32, row-major, `shape[ndim]`.  weight_g / weight_v pairs are folded by the engine. */
TTSB_API int ttsb_engine_set_tensor(ttsb_engine* e, const char* name, const float* data, const int64_t* shape, int ndim);
/* Fold weight-norm, repack for the kernels, upload.  Fails (TTSB_E_INVALID) naming the first missing tensor. */
TTSB_API int ttsb_engine_finalize(ttsb_engine* e);
TTSB_API int ttsb_engine_set_precision(ttsb_engine* e, int precision);
TTSB_API int ttsb_engine_get_precision(ttsb_engine* e);
/* keep per-stage intermediates of the next synthesis for ttsb_engine_read_tap (debug / parity localisation) */
TTSB_API int ttsb_engine_enable_taps(ttsb_engine* e, int on);
/* number of kernels this engine has launched since creation (bench.py's gpu_launches) */
TTSB_API int64_t ttsb_engine_launch_count(ttsb_engine* e);

/* high-water mark (bytes) of the engine's workspace during the last synthesis / vocoder step.  The workspace is a stack allocator:
 * stage buffers are handed back as soon as the launches that use them are queued (one stream), so the mark is what a batch really
 * needs - about 0.3 MB per decoder frame in the 16-bit operand modes */
TTSB_API int ttsb_engine_workspace_peak(ttsb_engine* e, int64_t* bytes);

/* device time (ms) of the last synthesis / vocoder step: CUDA events on the engine's stream from the first launch of
 * begin() (or vocoder()) to the last operation of finish(), host gaps between the two phases included */
TTSB_API int ttsb_engine_last_step_ms(ttsb_engine* e, float* ms);
/* per-kernel-category CUDA-event timing of the next steps (bench.py's roofline): category 0 tcgen05 conv, 1 fp32 conv,
 * 2 LSTM recurrence, 3 iSTFT, 4 upsample-add (elementwise), 5 length-regulation gather, 6 attention, 7 LayerNorm family (ALBERT
 * embedding LN, LN, AdaLayerNorm), 8 harmonic source (phase scan + SineGen/STFT), 9 fp32 row projections (style -> AdaIN scales
 * and shifts, F0_proj / N_proj).  Results are those of the last step. */
TTSB_API int ttsb_engine_profile(ttsb_engine* e, int on);
TTSB_API int ttsb_engine_profile_read(ttsb_engine* e, int category, double* ms, double* flops, double* bytes, int64_t* launches);

/* Threading: a ttsb_engine owns one CUDA stream and one workspace, and begin() ... finish() is a pair with state in between
 * (the host needs pred_dur to size the frame side).  Calls on ONE engine must not overlap: serialise them (the Python Engine
 * holds a lock across the pair) or create one engine per thread.  Different engines are independent.  A refused call
 * (TTSB_E_TOO_LONG, TTSB_E_INVALID, ...) leaves the engine usable for the next one. */
/* ---- the hot path, two-phase because the number of frames is data-dependent ---------------------------
 * replaces: BatchKModel.forward_batch (core/kmodel.py:7-150) and the inherited KModel.forward
 *           (core/pipeline.py:204); ids are already mapped through the vocab and wrapped in 0 ... 0
 *           (kmodel.py:36-42 stays in Python).
 *
 * begin():  kmodel.py:62-119 — ALBERT, bert_encoder, DurationEncoder, predictor.lstm, duration_proj,
 *           sigmoid-sum, /speed, round-half-even, clamp(min 1), alignment.
 *   ids            host int32 [token_offsets[B]]
 *   token_offsets  host int32 [B+1]
 *   ref_s          host fp32 [B,256]   (voice-pack row: [:128] decoder style, [128:] prosody style; kmodel.py:83-84)
 *   speeds         host fp32 [B]
 *   padded_lstm    != 0: reproduce forward_batch's un-packed predictor.lstm over the zero-padded batch
 *                  (kmodel.py:95); 0: per-utterance semantics of KModel.forward
 *   forced_dur     host int32 [sum T] or NULL: if given, these durations replace the predicted ones
 *                  (vocoder/decoder experiments; the predicted ones are still returned)
 * outputs (host, caller-owned):
 *   pred_dur       int32 [sum T]    — bit-exact contract
 *   duration       fp32 [sum T] or NULL — pre-round value (margin-to-.5 diagnostics)
 *   frame_offsets  int32 [B+1]      — utterance b owns frames [frame_offsets[b], frame_offsets[b+1]) and
 *                                     audio samples 600x that
 */
TTSB_API int ttsb_synthesize_begin(ttsb_engine* e, const int32_t* ids, const int32_t* token_offsets, int B, const float* ref_s,
                          const float* speeds, int padded_lstm, const int32_t* forced_dur, int32_t* pred_dur, float* duration,
                          int32_t* frame_offsets);

/* Voice packs resident on the device.  The reference keeps every loaded voice pack ([510,1,256] fp32) in a host
 * dict (core/pipeline.py:78-109 load_voice) and, per request, picks the row for the utterance's phoneme count and stacks the rows of a
 * batch into ref_s (pipeline.py:596-603) - one host gather + one host->device copy per batch.  Here a pack is uploaded once into a
 * slot; begin_voices() takes (slot, row) per utterance and gathers the style rows on the device.
 *   voice_pack_set: pack = host fp32 [rows,256] (the pack with its singleton dimension squeezed), copied; an existing pack in the slot
 *                   is replaced; pack == NULL empties the slot.  slot in [0, 4096).  Not to be called between begin() and finish().
 *   begin_voices:   as ttsb_synthesize_begin with voice_slot / voice_row (host int32 [B]) in place of ref_s; row = the index the
 *                   reference uses, len(phonemes) - 1, clamped by the caller to the pack (pipeline.py:599-601).  Results are
 *                   bit-identical to begin() with the same rows passed as ref_s.
 */
TTSB_API int ttsb_engine_voice_pack_set(ttsb_engine* e, int slot, const float* pack, int rows);
TTSB_API int ttsb_synthesize_begin_voices(ttsb_engine* e, const int32_t* ids, const int32_t* token_offsets, int B, const int32_t* voice_slot,
                                 const int32_t* voice_row, const float* speeds, int padded_lstm, const int32_t* forced_dur,
                                 int32_t* pred_dur, float* duration, int32_t* frame_offsets);

/* finish(): kmodel.py:120-143 — length regulation, F0Ntrain, text_encoder, decoder + iSTFTNet generator.
 *   noise       fp32 [600*sum F, 9] N(0,1) samples for SineGen (injected for parity), in `noise_mem`; or NULL:
 *               the engine draws them in-kernel (Philox).  The stream of an utterance is keyed by (`seed`, a hash of that
 *               utterance's own token ids, sample index within the utterance): the same utterance with the same `seed` draws
 *               the same noise wherever it sits in a batch and whatever else is in it; another `seed` gives another draw.
 *               (ttsb_vocoder has no token ids: there the stream is keyed by `seed` and the position in the batch's buffer.)
 *   indices     host int32 [sum F] or NULL — the alignment index map (token index of each frame; kmodel.py:116-119)
 *   audio       fp32 [600*sum F] in `audio_mem` (TTSB_MEM_HOST: the engine does the D2H copy; TTSB_MEM_DEVICE:
 *               written in place, valid when `stream` has drained)
 *   The call synchronises before returning when audio_mem == TTSB_MEM_HOST.                                         */
TTSB_API int ttsb_synthesize_finish(ttsb_engine* e, const float* noise, int noise_mem, uint64_t seed, int32_t* indices, float* audio,
                           int audio_mem);

/* ---- vocoder only (BASELINE config 4): kokoro Generator.forward ------------------------------------------
 *   x dev/host fp32 channels-last [sum L, 512] (generator input frames), f0 [sum L], style [B,128], l_offsets host [B+1]
 *   noise [300*sum L, 9] or NULL, audio [300*sum L]                                                          */
TTSB_API int ttsb_vocoder(ttsb_engine* e, const float* x, const float* f0, const float* style, const int32_t* l_offsets, int B, int in_mem,
                 const float* noise, uint64_t seed, float* audio, int audio_mem);

/* ---- taps -------------------------------------------------------------------------------------------------
 * read an intermediate of the last synthesis (enable_taps first): fp32 row-major [rows, cols] channels-last.
 * call with out == NULL to query the shape. */
TTSB_API int ttsb_engine_read_tap(ttsb_engine* e, const char* name, float* out, int64_t* rows, int64_t* cols);

/* ---- single kernels, for unit parity and roofline measurement (device pointers, engine's stream) ------------
 * each returns the CUDA-event time of the launch(es) in *ms when ms != NULL (the call then synchronises). */

/* fused conv_post epilogue + iSTFT(20,5) overlap-add (kokoro TorchSTFT.inverse on exp/sin of conv_post):
 *   cp dev fp32 [sum frames, ld >= 24, ld % 4 == 0] (cols 0..10 log-magnitude, 11..21 phase argument), frames_b = 120 F_b + 1,
 *   audio dev fp32 [600 sum F]                                                                                  */
TTSB_API int ttsb_op_istft(ttsb_engine* e, const float* cp, int ld, const int32_t* frame_counts_F, int B, float* audio, float* ms);
/* AdaIN1d + Snake over channels-last rows: y = v + sin^2(alpha v)/alpha, v = (1+g)(x-mean)rstd + b; statistics
 * per (utterance, channel) computed by the engine.  x,y dev fp32 [sum rows, C]; gamma_beta dev [B, 2C]; alpha dev [C] */
TTSB_API int ttsb_op_adain_snake(ttsb_engine* e, const float* x, float* y, int C, const int32_t* row_offsets, int B, const float* gamma_beta,
                        const float* alpha, float* ms_stats, float* ms_apply);
/* duration -> alignment index map -> gather (kmodel.py:116-139 without the one-hot matmul):
 *   pred_dur dev int32 [sum T]; src dev fp32 [sum T, C]; indices dev int32 [sum F]; dst dev fp32 [sum F, C]      */
TTSB_API int ttsb_op_length_regulate(ttsb_engine* e, const int32_t* pred_dur, const int32_t* token_offsets, int B, const float* src, int C,
                            int32_t* indices, float* dst, int32_t* frame_offsets_host, float* ms);
/* one Conv1d over channels-last rows: w host fp32 [Cout, Cin, k] (torch layout), bias host [Cout] or NULL;
 * x dev [sum rows, Cin], y dev [sum rows, Cout]; precision as above                                          */
TTSB_API int ttsb_op_conv1d(ttsb_engine* e, const float* x, const float* w, const float* bias, float* y, int Cin, int Cout, int k, int dilation,
                   const int32_t* row_offsets, int B, int precision, int reps, float* ms);
/* bidirectional LSTM (hidden 256/direction, PyTorch gate order): x dev [sum rows, In]; host weights in torch layout
 * w_ih [2][1024, In], w_hh [2][1024,256], b_ih/b_hh [2][1024]; y dev [sum rows, 512]                                */
TTSB_API int ttsb_op_bilstm(ttsb_engine* e, const float* x, int In, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                   const int32_t* row_offsets, int B, float* y, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* TTSB200_H_ */
