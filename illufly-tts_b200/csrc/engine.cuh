// Engine internals: weight store, ragged row spaces, workspace arena, graph runners.
// The C ABI in include/ttsb200.h is implemented in engine_api.cu on top of this.
#pragma once
#include "common.cuh"

#include <map>
#include <string>
#include <vector>

namespace ttsb {

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
  int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
};

// one Conv1d / Linear / polyphase ConvTranspose1d, packed for both conv kernels
struct ConvW {
  float* w = nullptr; int ldw = 0;        // fp32 [k][Cin][ldw]
  float* bias = nullptr;                  // [Cout] or null
  uint16_t* wb = nullptr;                 // bf16 UMMA pack (see conv_umma.cu), or null
  uint16_t* wh = nullptr;                 // fp16 UMMA pack, or null
  uint16_t* ws = nullptr;                 // split-fp16 UMMA pack (hi | lo), or null: fp32-equivalent tensor-core path
  uint16_t* wab = nullptr, *wah = nullptr; // bf16 / fp16 packs in conv_umma_wa.cu's stage order (256-channel resblock layers), or null
  int cin_pad = 0, cout_pad = 0;
  int cin_pad_s = 0, cout_pad_s = 0;
  float ws_scale = 0.f;                   // power of two the split pack's weights were scaled by (chosen from max|w| of the layer)
  int Cin = 0, Cout = 0, k = 1;
};
struct LstmW {
  ConvW ih;                               // [In] -> 2048 (fwd i,f,g,o | rev i,f,g,o), bias = b_ih + b_hh
  float* whh = nullptr;                   // [2][1024][256] (lstm_cluster_kernel: rows contiguous)
};
struct AdainRef { int off = 0; int C = 0; };  // slice of the per-utterance gamma|beta row
struct ResBlk1dW {                        // kokoro AdainResBlk1d
  ConvW conv1, conv2, sc;
  bool learned_sc = false, up = false;
  float* pool_w = nullptr; float* pool_b = nullptr;
  AdainRef n1, n2;
  int cin = 0, cout = 0;
};
struct ResBlock1W {                       // kokoro AdaINResBlock1
  ConvW c1[3], c2[3];
  AdainRef a1[3], a2[3];
  float* alpha1[3] = {nullptr, nullptr, nullptr};
  float* alpha2[3] = {nullptr, nullptr, nullptr};
  int C = 0, k = 0, dil[3] = {1, 3, 5};
};

// a ragged row space: utterance b owns rows [off[b], off[b]+len[b])
struct Space {
  std::vector<int> off, len;
  int total = 0, maxlen = 0;
  int n() const { return (int)len.size(); }
  void build(const std::vector<int>& lens) {
    len = lens; off.resize(lens.size()); total = 0; maxlen = 0;
    for (size_t i = 0; i < lens.size(); ++i) { off[i] = total; total += lens[i]; if (lens[i] > maxlen) maxlen = lens[i]; }
  }
};
struct SegTab {                           // device tables for one (in space -> out space) op
  const Seg* segs = nullptr; const Tile* tiles = nullptr;
  int nseg = 0, ntiles = 0, max_out = 0, max_in = 0, out_rows = 0, in_rows = 0;
};

struct Tap { float* d = nullptr; int64_t rows = 0, cols = 0; };

// per-category device time of the last step, from CUDA events recorded around the launches on the engine's stream
enum ProfCat : int { PROF_CONV_TENSOR = 0, PROF_CONV_SIMT = 1, PROF_LSTM = 2, PROF_ISTFT = 3, PROF_ELEMENTWISE = 4, PROF_GATHER = 5, PROF_ATTENTION = 6, PROF_LAYERNORM = 7,
                     PROF_SOURCE = 8, PROF_PROJ = 9 /* style projections, F0_proj / N_proj: fp32 row kernels */, PROF_NCAT = 10 };
struct ProfRec { int cat; double flops, bytes; };
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> events;
  std::vector<ProfRec> recs;
  double ms[PROF_NCAT] = {0}, flops[PROF_NCAT] = {0}, bytes[PROF_NCAT] = {0};
  long long launches[PROF_NCAT] = {0};
  void begin(cudaStream_t s, int cat, double fl, double by) {
    if (!on) return;
    const size_t i = recs.size();
    while (events.size() < 2 * (i + 1)) { cudaEvent_t ev; if (cudaEventCreate(&ev) != cudaSuccess) { on = false; return; } events.push_back(ev); }
    recs.push_back({cat, fl, by});
    cudaEventRecord(events[2 * i], s);
  }
  void end(cudaStream_t s) {
    if (!on || recs.empty()) return;
    cudaEventRecord(events[2 * (recs.size() - 1) + 1], s);
  }
  void reset() { recs.clear(); for (int c = 0; c < PROF_NCAT; ++c) { ms[c] = flops[c] = bytes[c] = 0; launches[c] = 0; } }
  void collect() {   // call after the stream has drained
    for (size_t i = 0; i < recs.size(); ++i) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, events[2 * i], events[2 * i + 1]) != cudaSuccess) { (void)cudaGetLastError(); continue; }
      ms[recs[i].cat] += t; flops[recs[i].cat] += recs[i].flops; bytes[recs[i].cat] += recs[i].bytes; launches[recs[i].cat] += 1;
    }
    recs.clear();
  }
  void release() { for (auto ev : events) cudaEventDestroy(ev); events.clear(); }
};

// Workspace of one engine: a stack allocator over a few large device allocations.  Everything of a step runs on ONE stream, so a
// buffer may be handed out again as soon as the launches that use it are QUEUED: release_to(mark) pops everything allocated since
// mark() (stream order keeps the earlier kernels ahead of the later writers).  Scopes nest; the peak is what a step really needs.
class Arena {
 public:
  struct Mark { size_t chunk, off; };
  ~Arena() { release(); }
  void release();
  void reset();                           // start a new synthesis (may consolidate chunks)
  void* alloc(size_t bytes);              // 256-byte aligned; nullptr on failure
  Mark mark() const { return {cur_, chunks_.empty() ? 0 : chunks_[cur_].off}; }
  void release_to(const Mark& m);
  size_t used() const { return total_used_; }
  size_t peak() const { return peak_; }   // high-water mark since the last reset()
 private:
  struct Chunk { char* p; size_t cap, off; };
  std::vector<Chunk> chunks_;
  size_t cur_ = 0;                        // chunk the next allocation is tried in first
  size_t total_used_ = 0, peak_ = 0;
};

struct Engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;               // independent branches of the graph (the N branch of F0Ntrain) run here, forked / joined by events
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_sync = nullptr, ev_step0 = nullptr, ev_step1 = nullptr;
  Profiler prof;
  int precision = 0;
  bool taps_on = false;
  cudaEvent_t ev_gap0 = nullptr, ev_gap1 = nullptr;   // TTSB_GAP_PROF=1: GPU idle time between the two halves of a step
  std::vector<unsigned long long> utt_key;   // per utterance of the step in flight: hash of its token ids (begin -> finish)
  // voice packs resident in device memory: slot -> [rows, 256] fp32; begin() can take (slot, row) per utterance
  std::vector<float*> voice_pack;
  std::vector<int> voice_pack_rows;
  bool finalized = false;
  std::map<std::string, HostTensor> host;       // raw checkpoint tensors until finalize()
  std::vector<void*> dev_allocs;                // weights
  std::map<std::string, Tap> taps;
  Arena arena;
  char* pinned = nullptr; size_t pinned_cap = 0;
  char* pinned_out = nullptr; size_t pinned_out_cap = 0;   // begin(): durations / totals / range counter come back through pinned memory
  unsigned int* d_sat = nullptr;          // split-fp16 operand range counter (ConvArgs::sat), zeroed at the start of a step

  // ---- weights ----
  float *emb_word = nullptr, *emb_pos = nullptr, *emb_type = nullptr, *emb_ln_g = nullptr, *emb_ln_b = nullptr;
  ConvW map_in, qkv, att_dense, ffn, ffn_out, bert_enc;
  float *att_ln_g = nullptr, *att_ln_b = nullptr, *full_ln_g = nullptr, *full_ln_b = nullptr;
  LstmW de_lstm[3], dur_lstm, shared_lstm, te_lstm;
  AdainRef de_aln[3];
  ConvW dur_proj;
  ConvW style_pred, style_dec;                  // 128 -> all AdaIN / AdaLayerNorm gamma|beta rows of one style half
  ResBlk1dW f0_blk[3], n_blk[3];
  ConvW f0_proj, n_proj;
  float* te_emb = nullptr;
  ConvW te_cnn[3]; float* te_ln_g[3] = {nullptr, nullptr, nullptr}; float* te_ln_b[3] = {nullptr, nullptr, nullptr};
  ResBlk1dW enc_blk, dec_blk[4];
  ConvW asr_res;
  float *f0conv_w = nullptr, *f0conv_b = nullptr, *nconv_w = nullptr, *nconv_b = nullptr;
  float *src_lin_w = nullptr, *src_lin_b = nullptr;
  ConvW noise_conv[2], ups[2], conv_post;
  ConvW noise_conv0r;                           // noise_convs[0] (k12, stride 6) as a k3 conv over rows regrouped 6 at a time
  ResBlock1W noise_res[2], resblk[6];

  // ---- state between begin() and finish() ----
  bool begun = false;
  int B = 0;
  Space spT, spF;
  float* d_refs = nullptr;      // [B,256]
  float* d_X = nullptr;         // DurationEncoder output d [sum T, 640]
  float* d_ten = nullptr;       // text_encoder output [sum T, 512]
  int* d_pred_dur = nullptr;    // [sum T]
  int* d_cum = nullptr;         // inclusive scan per utterance
  float* d_gbP = nullptr;       // [B, style_pred.Cout]
  SegTab tabT;
  int* d_row2segT = nullptr;
};

// implemented in engine.cu
int engine_finalize(Engine& e);
int engine_begin(Engine& e, const int32_t* ids, const int32_t* token_offsets, int B, const float* ref_s, const float* speeds,
                 int padded_lstm, const int32_t* forced_dur, int32_t* pred_dur, float* duration, int32_t* frame_offsets,
                 const int32_t* voice_slot = nullptr, const int32_t* voice_row = nullptr);   // (slot, row) of a resident voice pack instead of ref_s
int engine_finish(Engine& e, const float* noise, int noise_mem, uint64_t seed, int32_t* indices, float* audio, int audio_mem);
int engine_vocoder(Engine& e, const float* x, const float* f0, const float* style, const int32_t* l_offsets, int B, int in_mem,
                   const float* noise, uint64_t seed, float* audio, int audio_mem);
void set_error(const char* fmt, ...);

}  // namespace ttsb
