// The synthesis graph of illufly-tts' engine (BatchKModel.forward_batch, /root/reference/src/illufly_tts/core/kmodel.py:62-143;
// module maths = kokoro 0.9.4, restated in SURVEY.md Appendix A) scheduled as sm_100a kernels over ragged channels-last rows.
// No CPU arithmetic here besides the one-off weight fold/repack at load time and the segment tables.
#include "engine.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../../include/ttsb200.h"

namespace ttsb {

// bounds that keep every row / sample count of a step inside int32 (600 samples and 120 stage-1 rows per frame)
constexpr int MAX_DUR_PER_TOKEN = 1 << 16;
constexpr long long MAX_FRAMES_PER_CALL = 1000000;     // 6.9 hours of audio; the workspace (~0.6 MB per frame) gives out long before

long long g_launches = 0;
bool pdl_enabled() {
  static const bool on = getenv("TTSB_NO_PDL") == nullptr;
  return on;
}
static thread_local char g_err[1024] = "";
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#define CK(expr)                                                                                        \
  do {                                                                                                  \
    cudaError_t ck_e__ = (expr);                                                                        \
    if (ck_e__ != cudaSuccess) {                                                                        \
      set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(ck_e__), __FILE__, __LINE__, #expr);     \
      return TTSB_E_CUDA;                                                                               \
    }                                                                                                   \
  } while (0)
#define RC(expr)                                  \
  do {                                            \
    int rc__ = (expr);                            \
    if (rc__ != TTSB_OK) return rc__;             \
  } while (0)

// ======================================================================================================
// arena
// ======================================================================================================
void Arena::release() {
  for (auto& c : chunks_) cudaFree(c.p);
  chunks_.clear();
  cur_ = 0;
  total_used_ = 0;
}
void Arena::reset() {
  if (chunks_.size() > 1) {
    // the last step spilled into extra chunks: one allocation of the step's high-water mark (rounded up) replaces them
    size_t cap = ((peak_ + (peak_ >> 3)) + (size_t(256) << 20) - 1) & ~((size_t(256) << 20) - 1);
    release();
    char* p = nullptr;
    if (cudaMalloc(&p, cap) == cudaSuccess) chunks_.push_back({p, cap, 0});
    else (void)cudaGetLastError();
  }
  for (auto& c : chunks_) c.off = 0;
  cur_ = 0;
  total_used_ = 0;
  peak_ = 0;
}
void* Arena::alloc(size_t bytes) {
  bytes = (bytes + 255) & ~size_t(255);
  if (bytes == 0) bytes = 256;
  for (size_t i = cur_; i < chunks_.size(); ++i) {
    Chunk& c = chunks_[i];
    if (i > cur_ && c.off != 0) break;                 // chunks past the current one are empty by construction
    if (c.off + bytes <= c.cap) {
      void* r = c.p + c.off;
      // bytes skipped at the end of the chunks in between count as used: release_to() gives them back with the chunk
      for (size_t j = cur_; j < i; ++j) { total_used_ += chunks_[j].cap - chunks_[j].off; chunks_[j].off = chunks_[j].cap; }
      c.off += bytes; total_used_ += bytes; cur_ = i;
      if (total_used_ > peak_) peak_ = total_used_;
      return r;
    }
  }
  size_t cap = std::max(bytes, size_t(256) << 20);
  char* p = nullptr;
  if (cudaMalloc(&p, cap) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  for (size_t j = cur_; j < chunks_.size(); ++j) { total_used_ += chunks_[j].cap - chunks_[j].off; chunks_[j].off = chunks_[j].cap; }
  chunks_.push_back({p, cap, bytes});
  cur_ = chunks_.size() - 1;
  total_used_ += bytes;
  if (total_used_ > peak_) peak_ = total_used_;
  return p;
}
void Arena::release_to(const Mark& m) {
  if (chunks_.empty() || m.chunk >= chunks_.size()) return;
  for (size_t i = m.chunk + 1; i < chunks_.size(); ++i) { total_used_ -= chunks_[i].off; chunks_[i].off = 0; }
  total_used_ -= chunks_[m.chunk].off - m.off;
  chunks_[m.chunk].off = m.off;
  cur_ = m.chunk;
}

// ======================================================================================================
// weight fold / repack (host, once)
// ======================================================================================================
namespace {

struct Loader {
  Engine& e;
  std::string missing;
  explicit Loader(Engine& en) : e(en) {}

  const HostTensor* get(const std::string& name, bool required = true) {
    auto it = e.host.find(name);
    if (it == e.host.end()) {
      if (required && missing.empty()) missing = name;
      return nullptr;
    }
    return &it->second;
  }
  template <typename T>
  T* upload(const std::vector<T>& v) {
    T* d = nullptr;
    size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    if (cudaMalloc(&d, bytes) != cudaSuccess) { if (missing.empty()) missing = "<cudaMalloc failed>"; return nullptr; }
    e.dev_allocs.push_back(d);
    if (!v.empty() && cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
      if (missing.empty()) missing = "<cudaMemcpy failed>";
    }
    return d;
  }
  float* upload_named(const std::string& name) {
    const HostTensor* t = get(name);
    if (!t) return nullptr;
    return upload(t->data);
  }
  // plain or weight-normed (old style weight_g/weight_v, norm over all dims but 0) weight, as a dense host tensor
  bool weight(const std::string& prefix, HostTensor& out) {
    if (const HostTensor* w = get(prefix + ".weight", false)) { out = *w; return true; }
    const HostTensor* g = get(prefix + ".weight_g");
    const HostTensor* v = get(prefix + ".weight_v");
    if (!g || !v) return false;
    out.shape = v->shape;
    out.data.resize(v->data.size());
    const int64_t n0 = v->shape[0];
    const int64_t inner = v->numel() / n0;
    for (int64_t i = 0; i < n0; ++i) {
      double ss = 0.0;
      for (int64_t j = 0; j < inner; ++j) { const float x = v->data[i * inner + j]; ss += (double)x * (double)x; }
      // torch: w = v * (g / norm(v)); norm is computed in fp32
      const float nrm = (float)sqrt(ss);
      const float sc = g->data[i] / nrm;
      for (int64_t j = 0; j < inner; ++j) out.data[i * inner + j] = v->data[i * inner + j] * sc;
    }
    return true;
  }
  // torch Conv1d weight [Cout, Cin, k] (or Linear [Cout, Cin]) -> ConvW
  // want: bit 0 = 16-bit tensor-core packs (bf16 + fp16), bit 1 = split-fp16 pack (fp32-equivalent), bit 2 = weights-as-A packs (256 channels)
  void pack_conv(const HostTensor& w, const std::vector<float>* bias, ConvW& c, int want);
  bool conv(const std::string& prefix, ConvW& c, bool has_bias = true, int want = 0) {
    HostTensor w;
    if (!weight(prefix, w)) return false;
    const HostTensor* b = has_bias ? get(prefix + ".bias") : nullptr;
    if (has_bias && !b) return false;
    pack_conv(w, b ? &b->data : nullptr, c, want);
    return true;
  }
  // Conv1d whose channel counts are padded with zero weights / zero bias up to (cout_to, cin_to).  For the two 22-channel layers of
  // the generator (conv_post 128 -> 22, noise_convs[1] 22 -> 128): both buffers already have 24 floats per row with the last two
  // zero, and 22 is not a multiple of 4 (8), which sent the epilogue (the producers) of the tensor-core kernel down its scalar
  // ragged-channel path - 8.5 k cycles per tile instead of 1.2 k (TTSB_UMMA_PROF).  The extra output columns are exact zeros.
  bool conv_padded(const std::string& prefix, ConvW& c, int cout_to, int cin_to, int want) {
    HostTensor w;
    if (!weight(prefix, w)) return false;
    const HostTensor* b = get(prefix + ".bias");
    if (!b) return false;
    const int Co = (int)w.shape[0], Ci = (int)w.shape[1], k = w.shape.size() > 2 ? (int)w.shape[2] : 1;
    const int Co2 = std::max(Co, cout_to), Ci2 = std::max(Ci, cin_to);
    HostTensor wp;
    wp.shape = {Co2, Ci2, k};
    wp.data.assign((size_t)Co2 * Ci2 * k, 0.f);
    for (int co = 0; co < Co; ++co)
      for (int ci = 0; ci < Ci; ++ci)
        for (int j = 0; j < k; ++j) wp.data[((size_t)co * Ci2 + ci) * k + j] = w.data[((size_t)co * Ci + ci) * k + j];
    std::vector<float> bp((size_t)Co2, 0.f);
    for (int co = 0; co < Co && co < (int)b->data.size(); ++co) bp[co] = b->data[co];
    pack_conv(wp, &bp, c, want);
    return true;
  }
  bool lstm(const std::string& prefix, LstmW& l);
  bool resblk1d(const std::string& prefix, ResBlk1dW& r, int cin, int cout, bool up, std::vector<float>& style_w, std::vector<float>& style_b,
                int want);
  bool resblock1(const std::string& prefix, ResBlock1W& r, int C, int k, std::vector<float>& style_w, std::vector<float>& style_b);
  // append one AdaIN / AdaLayerNorm fc (Linear 128 -> n) to a concatenated style projection; returns its column offset
  bool style_fc(const std::string& prefix, int n, std::vector<float>& style_w, std::vector<float>& style_b, AdainRef& ref, int C) {
    const HostTensor* w = get(prefix + ".weight");
    const HostTensor* b = get(prefix + ".bias");
    if (!w || !b) return false;
    if (w->numel() != (int64_t)n * 128 || b->numel() != n) { if (missing.empty()) missing = prefix + " (shape)"; return false; }
    ref.off = (int)style_b.size();
    ref.C = C;
    style_w.insert(style_w.end(), w->data.begin(), w->data.end());   // rows = outputs, [n,128]
    style_b.insert(style_b.end(), b->data.begin(), b->data.end());
    while (style_b.size() & 3) { style_b.push_back(0.f); style_w.insert(style_w.end(), 128, 0.f); }
    return true;
  }
};

void Loader::pack_conv(const HostTensor& w, const std::vector<float>* bias, ConvW& c, int want) {
  const int Cout = (int)w.shape[0], Cin = (int)w.shape[1], k = w.shape.size() > 2 ? (int)w.shape[2] : 1;
  c.Cin = Cin; c.Cout = Cout; c.k = k;
  c.ldw = (Cout + 3) & ~3;
  std::vector<float> p((size_t)k * Cin * c.ldw, 0.f);
  for (int co = 0; co < Cout; ++co)
    for (int ci = 0; ci < Cin; ++ci)
      for (int j = 0; j < k; ++j) p[((size_t)j * Cin + ci) * c.ldw + co] = w.data[((size_t)co * Cin + ci) * k + j];
  c.w = upload(p);
  c.bias = bias ? upload(*bias) : nullptr;
  std::vector<uint16_t> pb;
  if (want & 1) {
    pack_umma_weights(p.data(), k, Cin, Cout, c.ldw, 0, pb, c.cin_pad, c.cout_pad);
    c.wb = upload(pb);
    pack_umma_weights(p.data(), k, Cin, Cout, c.ldw, 1, pb, c.cin_pad, c.cout_pad);
    c.wh = upload(pb);
  }
  if (want & 2) {
    pack_umma_weights(p.data(), k, Cin, Cout, c.ldw, 2, pb, c.cin_pad_s, c.cout_pad_s, &c.ws_scale);
    c.ws = upload(pb);
  }
  if ((want & 4) && Cin == Cout && Cin == 256) {      // 256-channel resblock layers: the weights-as-A kernel's own stage order
    pack_umma_wa_weights(p.data(), k, Cin, c.ldw, 0, pb);
    c.wab = upload(pb);
    pack_umma_wa_weights(p.data(), k, Cin, c.ldw, 1, pb);
    c.wah = upload(pb);
  }
}

bool Loader::lstm(const std::string& prefix, LstmW& l) {
  const char* sfx[2] = {"", "_reverse"};
  const HostTensor* wih[2]; const HostTensor* whh[2]; const HostTensor* bih[2]; const HostTensor* bhh[2];
  for (int d = 0; d < 2; ++d) {
    wih[d] = get(prefix + ".weight_ih_l0" + sfx[d]);
    whh[d] = get(prefix + ".weight_hh_l0" + sfx[d]);
    bih[d] = get(prefix + ".bias_ih_l0" + sfx[d]);
    bhh[d] = get(prefix + ".bias_hh_l0" + sfx[d]);
    if (!wih[d] || !whh[d] || !bih[d] || !bhh[d]) return false;
  }
  const int In = (int)wih[0]->shape[1];
  HostTensor w; w.shape = {2048, In}; w.data.resize((size_t)2048 * In);
  std::vector<float> bias(2048);
  for (int d = 0; d < 2; ++d) {
    memcpy(&w.data[(size_t)d * 1024 * In], wih[d]->data.data(), sizeof(float) * 1024 * In);
    for (int g = 0; g < 1024; ++g) bias[d * 1024 + g] = bih[d]->data[g] + bhh[d]->data[g];
  }
  pack_conv(w, &bias, l.ih, 3);
  std::vector<float> o((size_t)2 * 1024 * 256);
  for (int d = 0; d < 2; ++d) memcpy(&o[(size_t)d * 1024 * 256], whh[d]->data.data(), sizeof(float) * 1024 * 256);
  l.whh = upload(o);
  return true;
}

bool Loader::resblk1d(const std::string& prefix, ResBlk1dW& r, int cin, int cout, bool up, std::vector<float>& sw, std::vector<float>& sb,
                      int want) {
  r.cin = cin; r.cout = cout; r.up = up; r.learned_sc = cin != cout;
  bool ok = conv(prefix + ".conv1", r.conv1, true, want) && conv(prefix + ".conv2", r.conv2, true, want);
  if (r.learned_sc) ok = ok && conv(prefix + ".conv1x1", r.sc, false, want);
  ok = ok && style_fc(prefix + ".norm1.fc", 2 * cin, sw, sb, r.n1, cin) && style_fc(prefix + ".norm2.fc", 2 * cout, sw, sb, r.n2, cout);
  if (up) {
    HostTensor w;
    if (!weight(prefix + ".pool", w)) return false;     // [C,1,3]
    const HostTensor* b = get(prefix + ".pool.bias");
    if (!b) return false;
    r.pool_w = upload(w.data);
    r.pool_b = upload(b->data);
  }
  return ok;
}

bool Loader::resblock1(const std::string& prefix, ResBlock1W& r, int C, int k, std::vector<float>& sw, std::vector<float>& sb) {
  r.C = C; r.k = k;
  bool ok = true;
  for (int m = 0; m < 3 && ok; ++m) {
    const std::string i = std::to_string(m);
    ok = ok && conv(prefix + ".convs1." + i, r.c1[m], true, 7) && conv(prefix + ".convs2." + i, r.c2[m], true, 7);
    ok = ok && style_fc(prefix + ".adain1." + i + ".fc", 2 * C, sw, sb, r.a1[m], C) && style_fc(prefix + ".adain2." + i + ".fc", 2 * C, sw, sb, r.a2[m], C);
    r.alpha1[m] = upload_named(prefix + ".alpha1." + i);
    r.alpha2[m] = upload_named(prefix + ".alpha2." + i);
    ok = ok && r.alpha1[m] && r.alpha2[m];
  }
  return ok;
}

// ConvTranspose1d(Cin, Cout, K = 2*stride, stride, padding = stride/2) as `stride` polyphase 2-tap convolutions:
//   Z[m, r*Cout + co] = sum_ci x[m-1, ci] W[ci, co, r + stride] + x[m, ci] W[ci, co, r]     (m = 0..L)
//   out[t] = Z_flat[t + stride/2]   with Z_flat = Z viewed as [(L+1)*stride, Cout]
bool pack_convT(Loader& ld, const std::string& prefix, int stride, ConvW& c) {
  HostTensor w;
  if (!ld.weight(prefix, w)) return false;            // [Cin, Cout, K]
  const HostTensor* b = ld.get(prefix + ".bias");
  if (!b) return false;
  const int Cin = (int)w.shape[0], Cout = (int)w.shape[1], K = (int)w.shape[2];
  if (K != 2 * stride) { if (ld.missing.empty()) ld.missing = prefix + " (kernel != 2*stride)"; return false; }
  HostTensor p; p.shape = {(int64_t)stride * Cout, Cin, 2}; p.data.resize((size_t)stride * Cout * Cin * 2);
  std::vector<float> bias((size_t)stride * Cout);
  for (int r = 0; r < stride; ++r)
    for (int co = 0; co < Cout; ++co) {
      bias[(size_t)r * Cout + co] = b->data[co];
      for (int ci = 0; ci < Cin; ++ci) {
        float* dst = &p.data[(((size_t)r * Cout + co) * Cin + ci) * 2];
        dst[0] = w.data[((size_t)ci * Cout + co) * K + r + stride];   // tap 0 reads x[m-1]
        dst[1] = w.data[((size_t)ci * Cout + co) * K + r];            // tap 1 reads x[m]
      }
    }
  ld.pack_conv(p, &bias, c, 3);
  return true;
}

}  // namespace

int engine_finalize(Engine& e) {
  if (e.finalized) return TTSB_OK;
  Loader ld(e);
  std::vector<float> swP, sbP, swD, sbD;   // concatenated style projections (prosody style / decoder style)
  bool ok = true;
  const std::string L = "bert.encoder.albert_layer_groups.0.albert_layers.0.";
  e.emb_word = ld.upload_named("bert.embeddings.word_embeddings.weight");
  e.emb_pos = ld.upload_named("bert.embeddings.position_embeddings.weight");
  if (const HostTensor* t = ld.get("bert.embeddings.token_type_embeddings.weight")) {
    std::vector<float> row0(t->data.begin(), t->data.begin() + 128);
    e.emb_type = ld.upload(row0);
  }
  e.emb_ln_g = ld.upload_named("bert.embeddings.LayerNorm.weight");
  e.emb_ln_b = ld.upload_named("bert.embeddings.LayerNorm.bias");
  ok = ok && ld.conv("bert.encoder.embedding_hidden_mapping_in", e.map_in, true, 2);
  {  // q | k | v fused into one 768 -> 2304 projection
    const HostTensor *wq = ld.get(L + "attention.query.weight"), *wk = ld.get(L + "attention.key.weight"), *wv = ld.get(L + "attention.value.weight");
    const HostTensor *bq = ld.get(L + "attention.query.bias"), *bk = ld.get(L + "attention.key.bias"), *bv = ld.get(L + "attention.value.bias");
    if (wq && wk && wv && bq && bk && bv) {
      HostTensor w; w.shape = {2304, 768};
      w.data = wq->data; w.data.insert(w.data.end(), wk->data.begin(), wk->data.end()); w.data.insert(w.data.end(), wv->data.begin(), wv->data.end());
      std::vector<float> b = bq->data; b.insert(b.end(), bk->data.begin(), bk->data.end()); b.insert(b.end(), bv->data.begin(), bv->data.end());
      ld.pack_conv(w, &b, e.qkv, 2);
    } else ok = false;
  }
  ok = ok && ld.conv(L + "attention.dense", e.att_dense, true, 2) && ld.conv(L + "ffn", e.ffn, true, 2) && ld.conv(L + "ffn_output", e.ffn_out, true, 2);
  e.att_ln_g = ld.upload_named(L + "attention.LayerNorm.weight"); e.att_ln_b = ld.upload_named(L + "attention.LayerNorm.bias");
  e.full_ln_g = ld.upload_named(L + "full_layer_layer_norm.weight"); e.full_ln_b = ld.upload_named(L + "full_layer_layer_norm.bias");
  ok = ok && ld.conv("bert_encoder", e.bert_enc, true, 2);
  for (int l = 0; l < 3 && ok; ++l) {
    ok = ok && ld.lstm("predictor.text_encoder.lstms." + std::to_string(2 * l), e.de_lstm[l]);
    ok = ok && ld.style_fc("predictor.text_encoder.lstms." + std::to_string(2 * l + 1) + ".fc", 1024, swP, sbP, e.de_aln[l], 512);
  }
  ok = ok && ld.lstm("predictor.lstm", e.dur_lstm) && ld.lstm("predictor.shared", e.shared_lstm) && ld.lstm("text_encoder.lstm", e.te_lstm);
  ok = ok && ld.conv("predictor.duration_proj.linear_layer", e.dur_proj, true, 2);
  const int pc[3][2] = {{512, 512}, {512, 256}, {256, 256}};
  for (int i = 0; i < 3 && ok; ++i) {
    ok = ok && ld.resblk1d("predictor.F0." + std::to_string(i), e.f0_blk[i], pc[i][0], pc[i][1], i == 1, swP, sbP, 3);
    ok = ok && ld.resblk1d("predictor.N." + std::to_string(i), e.n_blk[i], pc[i][0], pc[i][1], i == 1, swP, sbP, 3);
  }
  ok = ok && ld.conv("predictor.F0_proj", e.f0_proj) && ld.conv("predictor.N_proj", e.n_proj);
  e.te_emb = ld.upload_named("text_encoder.embedding.weight");
  for (int l = 0; l < 3 && ok; ++l) {
    const std::string p = "text_encoder.cnn." + std::to_string(l);
    ok = ok && ld.conv(p + ".0", e.te_cnn[l], true, 2);
    e.te_ln_g[l] = ld.upload_named(p + ".1.gamma");
    e.te_ln_b[l] = ld.upload_named(p + ".1.beta");
  }
  ok = ok && ld.resblk1d("decoder.encode", e.enc_blk, 514, 1024, false, swD, sbD, 3);
  for (int i = 0; i < 4 && ok; ++i)
    ok = ok && ld.resblk1d("decoder.decode." + std::to_string(i), e.dec_blk[i], 1090, i == 3 ? 512 : 1024, i == 3, swD, sbD, 3);
  ok = ok && ld.conv("decoder.asr_res.0", e.asr_res, true, 3);
  {
    HostTensor w;
    if (ld.weight("decoder.F0_conv", w)) e.f0conv_w = ld.upload(w.data); else ok = false;
    if (ld.weight("decoder.N_conv", w)) e.nconv_w = ld.upload(w.data); else ok = false;
    e.f0conv_b = ld.upload_named("decoder.F0_conv.bias");
    e.nconv_b = ld.upload_named("decoder.N_conv.bias");
  }
  const std::string G = "decoder.generator.";
  e.src_lin_w = ld.upload_named(G + "m_source.l_linear.weight");
  e.src_lin_b = ld.upload_named(G + "m_source.l_linear.bias");
  ok = ok && ld.conv(G + "noise_convs.0", e.noise_conv[0], true, 0) && ld.conv_padded(G + "noise_convs.1", e.noise_conv[1], 0, 24, 3);
  if (ok) {
    // noise_convs[0]: Conv1d(22, 256, k12, stride 6, pad 3) over STFT frames == k3/pad1 conv over frames regrouped six at a time
    // ([blocks, 6 x 24] view of the channels-last har buffer): tap in {0,1,2}, slot i in 0..5  <->  original tap j = 6 tap + i - 3
    const HostTensor* w0 = ld.get(G + "noise_convs.0.weight");
    const HostTensor* b0 = ld.get(G + "noise_convs.0.bias");
    if (w0 && b0 && w0->shape.size() == 3 && w0->shape[1] == 22 && w0->shape[2] == 12) {
      const int Co = (int)w0->shape[0];
      HostTensor wr; wr.shape = {Co, 144, 3}; wr.data.assign((size_t)Co * 144 * 3, 0.f);
      for (int co = 0; co < Co; ++co)
        for (int tap = 0; tap < 3; ++tap)
          for (int i = 0; i < 6; ++i) {
            const int j = 6 * tap + i - 3;
            if (j < 0 || j >= 12) continue;
            for (int c = 0; c < 22; ++c) wr.data[((size_t)co * 144 + i * 24 + c) * 3 + tap] = w0->data[((size_t)co * 22 + c) * 12 + j];
          }
      ld.pack_conv(wr, &b0->data, e.noise_conv0r, 3);
    } else ok = false;
  }
  ok = ok && pack_convT(ld, G + "ups.0", 10, e.ups[0]) && pack_convT(ld, G + "ups.1", 6, e.ups[1]);
  ok = ok && ld.resblock1(G + "noise_res.0", e.noise_res[0], 256, 7, swD, sbD) && ld.resblock1(G + "noise_res.1", e.noise_res[1], 128, 11, swD, sbD);
  const int rk[3] = {3, 7, 11};
  for (int i = 0; i < 6 && ok; ++i) ok = ok && ld.resblock1(G + "resblocks." + std::to_string(i), e.resblk[i], i < 3 ? 256 : 128, rk[i % 3], swD, sbD);
  ok = ok && ld.conv_padded(G + "conv_post", e.conv_post, 24, 0, 3);
  if (ok) {
    HostTensor w;
    w.shape = {(int64_t)sbP.size(), 128}; w.data = swP; ld.pack_conv(w, &sbP, e.style_pred, 0);
    w.shape = {(int64_t)sbD.size(), 128}; w.data = swD; ld.pack_conv(w, &sbD, e.style_dec, 0);
  }
  if (!ok || !ld.missing.empty()) {
    set_error("finalize: missing or malformed checkpoint tensor '%s'", ld.missing.empty() ? "?" : ld.missing.c_str());
    return TTSB_E_INVALID;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { set_error("finalize: upload failed"); return TTSB_E_CUDA; }
  e.host.clear();
  e.finalized = true;
  return TTSB_OK;
}

// ======================================================================================================
// run-time helpers
// ======================================================================================================
namespace {

struct Run {
  Engine& e;
  cudaStream_t s;
  std::vector<char> stage;        // host staging of segment tables, uploaded in one copy
  char* d_tables = nullptr;
  size_t tables_cap = 0;
  double* stats_base = nullptr; size_t stats_cap = 0, stats_off = 0;
  uint8_t* xpack = nullptr; size_t xpack_bytes = 0;     // scratch of the pre-packed operand images (text-side GEMMs), reused call after call
  Run* parent = nullptr;          // a side-stream Run (independent branch of the graph) draws statistics slots from its parent's pool
  explicit Run(Engine& en) : e(en), s(en.stream) {}
  Run(Run& par, cudaStream_t side) : e(par.e), s(side), parent(&par) {}

  template <typename T>
  T* alloc(size_t n) {
    T* p = (T*)e.arena.alloc(n * sizeof(T));
    if (!p) set_error("workspace allocation of %zu bytes failed", n * sizeof(T));
    return p;
  }
  // workspace scopes (Arena): everything allocated since `m` is handed back; the launches that use it are already queued on
  // the engine's one stream, ahead of whatever reuses the memory.  The operand-image scratch may live in the released region.
  Arena::Mark mark() const { return e.arena.mark(); }
  void release_to(const Arena::Mark& m) {
    if (xpack && (m.chunk < xpack_at.chunk || (m.chunk == xpack_at.chunk && m.off <= xpack_at.off))) { xpack = nullptr; xpack_bytes = 0; }
    e.arena.release_to(m);
  }
  Arena::Mark xpack_at{0, 0};
  // ---- segment tables: reserve device space first, fill host staging, upload once -----------------
  int tables_begin(size_t cap) {
    d_tables = (char*)e.arena.alloc(cap);
    if (!d_tables) { set_error("workspace allocation failed (tables)"); return TTSB_E_NOMEM; }
    tables_cap = cap;
    stage.clear();
    return TTSB_OK;
  }
  template <typename T>
  const T* push(const std::vector<T>& v) {
    size_t off = (stage.size() + 15) & ~size_t(15);
    stage.resize(off + v.size() * sizeof(T));
    if (!v.empty()) memcpy(stage.data() + off, v.data(), v.size() * sizeof(T));
    return (const T*)(d_tables + off);
  }
  SegTab make_tab(const Space& in, const Space& out, int in_scale = 1, int out_scale = 1, bool want_tiles = true, int tile_rows = 128) {
    std::vector<Seg> segs(in.n());
    std::vector<Tile> tiles;
    SegTab t;
    for (int b = 0; b < in.n(); ++b) {
      segs[b] = {in.off[b] * in_scale, in.len[b] * in_scale, out.off[b] * out_scale, out.len[b] * out_scale};
      if (want_tiles)
        for (int t0 = 0; t0 < segs[b].out_len; t0 += tile_rows) tiles.push_back({b, t0});
    }
    t.segs = push(segs); t.tiles = push(tiles);
    t.nseg = in.n(); t.ntiles = (int)tiles.size();
    t.max_out = out.maxlen * out_scale; t.max_in = in.maxlen * in_scale;
    t.out_rows = out.total * out_scale; t.in_rows = in.total * in_scale;
    return t;
  }
  SegTab make_tab_segs(const std::vector<Seg>& segs, bool want_tiles = true, int tile_rows = 128) {
    std::vector<Tile> tiles;
    SegTab t;
    for (int b = 0; b < (int)segs.size(); ++b) {
      if (want_tiles)
        for (int t0 = 0; t0 < segs[b].out_len; t0 += tile_rows) tiles.push_back({b, t0});
      t.max_out = std::max(t.max_out, segs[b].out_len); t.max_in = std::max(t.max_in, segs[b].in_len);
      t.out_rows += segs[b].out_len; t.in_rows += segs[b].in_len;
    }
    t.segs = push(segs); t.tiles = push(tiles);
    t.nseg = (int)segs.size(); t.ntiles = (int)tiles.size();
    return t;
  }
  int tables_upload() {
    if (stage.size() > tables_cap) { set_error("internal: table staging overflow"); return TTSB_E_INVALID; }
    if (stage.size() > e.pinned_cap) {
      if (e.pinned) cudaFreeHost(e.pinned);
      e.pinned_cap = std::max(stage.size() * 2, size_t(1) << 20);
      CK(cudaMallocHost(&e.pinned, e.pinned_cap));
    }
    // the previous upload from this pinned buffer must have drained before it is overwritten
    CK(cudaEventSynchronize(e.ev_sync));
    memcpy(e.pinned, stage.data(), stage.size());
    CK(cudaMemcpyAsync(d_tables, e.pinned, stage.size(), cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(e.ev_sync, s));
    return TTSB_OK;
  }
  // ---- statistics slots (double sum, sum of squares per (utterance, channel)) ---------------------
  int stats_begin(int nseg, size_t channels_per_seg) {
    stats_cap = (size_t)nseg * channels_per_seg * 2;
    stats_base = alloc<double>(stats_cap);
    if (!stats_base) return TTSB_E_NOMEM;
    stats_off = 0;
    CK(cudaMemsetAsync(stats_base, 0, stats_cap * sizeof(double), s));
    return TTSB_OK;
  }
  double* stats(int nseg, int C) {
    if (parent) return parent->stats(nseg, C);
    const size_t n = (size_t)nseg * C * 2;
    if (stats_off + n > stats_cap) { set_error("internal: statistics pool exhausted"); return nullptr; }
    double* p = stats_base + stats_off;
    stats_off += n;
    return p;
  }
  // tap of a buffer whose utterances are not contiguous (row offsets / lengths given per utterance)
  int tap_ragged(const char* name, const float* d, int64_t cols, int64_t ld, const std::vector<int>& offs, const std::vector<int>& lens) {
    if (!e.taps_on) return TTSB_OK;
    int64_t rows = 0;
    for (int l : lens) rows += l;
    Tap& t = e.taps[name];
    if (t.d) cudaFree(t.d);
    t.d = nullptr; t.rows = rows; t.cols = cols;
    if (rows * cols == 0) return TTSB_OK;
    CK(cudaMalloc(&t.d, sizeof(float) * rows * cols));
    int64_t o = 0;
    for (size_t b = 0; b < lens.size(); ++b) {
      CK(cudaMemcpy2DAsync(t.d + o * cols, cols * sizeof(float), d + (int64_t)offs[b] * ld, ld * sizeof(float), cols * sizeof(float), lens[b],
                           cudaMemcpyDeviceToDevice, s));
      o += lens[b];
    }
    return TTSB_OK;
  }
  int tap(const char* name, const float* d, int64_t rows, int64_t cols, int64_t ld) {
    if (!e.taps_on) return TTSB_OK;
    Tap& t = e.taps[name];
    if (t.d) cudaFree(t.d);
    t.d = nullptr; t.rows = rows; t.cols = cols;
    if (rows * cols == 0) return TTSB_OK;
    CK(cudaMalloc(&t.d, sizeof(float) * rows * cols));
    CK(cudaMemcpy2DAsync(t.d, cols * sizeof(float), d, ld * sizeof(float), cols * sizeof(float), rows, cudaMemcpyDeviceToDevice, s));
    return TTSB_OK;
  }
};

// one convolution call; defaults = plain conv with bias
struct ConvCall {
  const ConvW* w = nullptr;
  const float* x = nullptr; int ldx = 0;
  float* y = nullptr; int ldy = 0;
  const float* res = nullptr; int ldr = 0;
  int dil = 1, pad = 0, stride = 1, in_shr = 0;
  int pro = PRO_NONE; float slope = 0.f;
  const double* pstats = nullptr; const float* gb = nullptr; int ld_gb = 0; AdainRef ad; const float* alpha = nullptr;
  float out_scale = 1.f; int accumulate = 0; int act = ACT_NONE; float act_slope = 0.f;
  double* ostats = nullptr;
  bool tensor = false;           // eligible for the tcgen05 path (16-bit operands) when the engine runs in a tensor precision
  bool exact = false;            // text side: tcgen05 with split-fp16 operands (fp32-equivalent) in tensor precisions, FFMA in FP32
  bool wa = false;               // 128-channel layer on the weights-as-A kernel (conv_umma_wa.cu); the table has 256-row tiles
  bool x_f16 = false, y_f16 = false, r_f16 = false;   // wa only: x / y / res are fp16 rows (the intermediates of a resblock)
};

// Which kernel a convolution runs on, by precision mode:
//   FP32 (the strict tier)  every tensor-eligible layer, text side and frame side, on tcgen05 with split-fp16 operands (fp32-class
//                           products, 3 MMAs per K step); the FFMA kernel only where the tcgen05 kernel does not support the shape
//   FP32_FFMA               the CUDA-core fp32 kernel everywhere (the arithmetic reference of the strict tier; tests)
//   BF16 / FP16             text side (c.exact) split-fp16, frame side (c.tensor) 16-bit operands
int run_conv(Run& r, const ConvCall& c, const SegTab& t) {
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  const int prec = r.e.precision;
  const bool f16 = prec == TTSB_PRECISION_FP16;
  const bool ffma_only = prec == TTSB_PRECISION_FP32_FFMA;
  const bool split = !ffma_only && c.w->ws != nullptr && (c.exact || (c.tensor && prec == TTSB_PRECISION_FP32));
  const bool half16 = !ffma_only && !split && c.tensor && !c.exact && (prec == TTSB_PRECISION_FP16 || prec == TTSB_PRECISION_BF16);
  a.x = c.x; a.ldx = c.ldx; a.w = c.w->w; a.ldw = c.w->ldw;
  if (split) { a.wb = c.w->ws; a.wfmt = 2; a.cin_pad = c.w->cin_pad_s; a.cout_pad = c.w->cout_pad_s; a.split_wscale = c.w->ws_scale; a.sat = r.e.d_sat; }
  else if (half16) { a.wb = f16 ? c.w->wh : c.w->wb; a.wfmt = f16 ? 1 : 0; a.cin_pad = c.w->cin_pad; a.cout_pad = c.w->cout_pad; }
  a.bias = c.w->bias; a.y = c.y; a.ldy = c.ldy; a.res = c.res; a.ldr = c.ldr;
  a.Cin = c.w->Cin; a.Cout = c.w->Cout; a.k = c.w->k; a.dil = c.dil; a.pad = c.pad; a.stride = c.stride; a.in_shr = c.in_shr;
  a.pro = c.pro; a.slope = c.slope;
  a.pstats = c.pstats; a.ld_pstats = c.ad.C;
  a.gb = c.gb ? c.gb + c.ad.off : nullptr; a.ld_gb = c.ld_gb; a.beta_off = c.ad.C;
  a.alpha = c.alpha;
  a.out_scale = c.out_scale; a.accumulate = c.accumulate; a.act = c.act; a.act_slope = c.act_slope;
  a.ostats = c.ostats; a.ld_ostats = c.w->Cout;
  a.segs = t.segs; a.tiles = t.tiles; a.ntiles = t.ntiles; a.tile_rows = 128;
  if (c.wa) {
    // the caller sized its buffers for this kernel (fp16 intermediate, 256-row tiles): no other kernel can take the call
    a.tile_rows = CONV_WA_TILE_ROWS; a.x_f16 = c.x_f16; a.y_f16 = c.y_f16; a.r_f16 = c.r_f16;
    if (half16 && a.Cin == 256) a.wb = f16 ? c.w->wah : c.w->wab;        // that kernel's own stage order for two channel halves
    if (!half16 || !conv_umma_wa_supported(a)) { set_error("internal: weights-as-A conv kernel refused Cin=%d Cout=%d k=%d dil=%d", a.Cin, a.Cout, a.k, a.dil); return TTSB_E_INVALID; }
    const double rows_wa = (double)t.out_rows;
    r.e.prof.begin(r.s, PROF_CONV_TENSOR, 2.0 * rows_wa * a.Cin * a.Cout * a.k,
                   rows_wa * a.Cout * ((a.y_f16 ? 2.0 : 4.0) + (a.res ? (a.r_f16 ? 2.0 : 4.0) : 0.0) + (a.accumulate ? 4.0 : 0.0)) +
                       (a.x_f16 ? 2.0 : 4.0) * (double)t.in_rows * a.Cin);
    CK(launch_conv_umma_wa(a, r.s));
    r.e.prof.end(r.s);
    return TTSB_OK;
  }
  if ((split || half16) && a.stride == 1) {
    // layers with several Cout tiles read pre-packed operand images (conv_umma.cu: pack_act_kernel); scratch sized for this call
    const int ks = split ? 64 : (a.cin_pad < 128 ? a.cin_pad : 128);
    const int bn = split ? 128 : 256;
    if (a.cout_pad > bn) {
      const size_t sr = (size_t)((128 + (a.k - 1) * a.dil) | 1);
      size_t slab = (size_t)(ks / 8) * sr * 16 * (split ? 2 : 1);
      slab = (slab + 127) & ~(size_t)127;
      const size_t need = (size_t)t.ntiles * (a.cin_pad / ks) * slab;
      if (need > r.xpack_bytes && need <= ((size_t)512 << 20)) {
        const size_t cap = need > ((size_t)32 << 20) ? need : ((size_t)32 << 20);
        const Arena::Mark at = r.mark();
        uint8_t* pnew = r.alloc<uint8_t>(cap);
        if (pnew) { r.xpack = pnew; r.xpack_bytes = cap; r.xpack_at = at; }
      }
      a.xpack = r.xpack; a.xpack_bytes = r.xpack_bytes;
    }
  }
  if (c.pro >= PRO_ADAIN_LRELU && (!c.pstats || !c.gb)) { set_error("internal: AdaIN prologue without statistics"); return TTSB_E_INVALID; }
  const double rows = (double)t.out_rows;
  const double fl = 2.0 * rows * a.Cin * a.Cout * a.k;
  const double by = 4.0 * rows * ((double)a.Cout * (1 + (a.res ? 1 : 0) + (a.accumulate ? 1 : 0))) + 4.0 * (double)t.in_rows * a.Cin;
  if ((split || half16) && conv_umma_supported(a)) {
    r.e.prof.begin(r.s, PROF_CONV_TENSOR, fl, by);
    CK(launch_conv_umma(a, r.s));
  } else {
    static const bool conv_log = getenv("TTSB_CONV_LOG") != nullptr;
    if (conv_log && (split || half16))
      fprintf(stderr, "ttsb: conv Cin=%d Cout=%d k=%d dil=%d ldx=%d ldy=%d in_shr=%d wfmt=%d falls back to the FFMA kernel\n", a.Cin, a.Cout, a.k, a.dil, a.ldx, a.ldy,
              a.in_shr, a.wfmt);
    a.wb = nullptr; a.wfmt = 0; a.sat = nullptr;
    r.e.prof.begin(r.s, PROF_CONV_SIMT, fl, by);
    CK(launch_conv_simt(a, r.s));
  }
  r.e.prof.end(r.s);
  return TTSB_OK;
}

// utterances sorted by step count and dealt LSTM_UB at a time to the clusters of lstm_cluster_kernel
std::vector<LstmGroup> make_lstm_groups(const std::vector<LstmSeg>& segs) {
  std::vector<int> order(segs.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return segs[a].len + segs[a].pre > segs[b].len + segs[b].pre; });
  std::vector<LstmGroup> g((segs.size() + LSTM_UB - 1) / LSTM_UB);
  for (auto& x : g) memset(&x, 0, sizeof(x));
  for (size_t i = 0; i < order.size(); ++i) {
    LstmGroup& x = g[i / LSTM_UB];
    const LstmSeg& sg = segs[order[i]];
    x.off[i % LSTM_UB] = sg.off; x.len[i % LSTM_UB] = sg.len; x.pre[i % LSTM_UB] = sg.pre;
  }
  return g;
}

// x W_ih -> gates, then the recurrence
int run_bilstm(Run& r, const LstmW& l, const float* x, int ldx, float* gx, float* y, int ldy, const SegTab& t, const LstmGroup* groups, int ngroups,
               bool exact) {
  ConvCall c; c.w = &l.ih; c.x = x; c.ldx = ldx; c.y = gx; c.ldy = 2048; c.exact = exact; c.tensor = !exact;
  RC(run_conv(r, c, t));
  r.e.prof.begin(r.s, PROF_LSTM, 2.0 * t.out_rows * 2 * 1024 * 256, 0.0);
  CK(launch_lstm_cluster(gx, l.whh, l.ih.bias, y, ldy, groups, ngroups, r.s));
  r.e.prof.end(r.s);
  return TTSB_OK;
}

// kokoro AdainResBlk1d.  x [in rows, ldx] with statistics xstats (ld = blk.cin); result -> y (ldy), statistics -> ystats.
// `tin` : table in->in rows, `tout`: out->out rows, `tio`: in->out rows (only for upsampling blocks).
int run_resblk1d(Run& r, const ResBlk1dW& b, const float* x, int ldx, const double* xstats, float* y, int ldy, double* ystats,
                 const float* gb, int ld_gb, const SegTab& tin, const SegTab& tout, const SegTab& tio, bool tensor) {
  const int nseg = tin.nseg;
  const float rs2 = 0.70710678118654752440f;   // torch.rsqrt(torch.tensor(2.)) in fp32
  const int out_rows = b.up ? tout.out_rows : tin.out_rows;
  float* r1 = r.alloc<float>((size_t)out_rows * b.cout);
  double* st1 = r.stats(nseg, b.cout);
  if (!r1 || !st1) return TTSB_E_NOMEM;
  const float* scp = x; int ldsc = ldx;
  ConvCall c1; c1.w = &b.conv1; c1.pad = 1; c1.y = r1; c1.ldy = b.cout; c1.ostats = st1; c1.tensor = tensor;
  if (!b.up) {
    c1.x = x; c1.ldx = ldx; c1.pro = PRO_ADAIN_LRELU; c1.slope = 0.2f; c1.pstats = xstats; c1.gb = gb; c1.ld_gb = ld_gb; c1.ad = b.n1;
    RC(run_conv(r, c1, tin));
  } else {
    const int ldp = (b.cin + 3) & ~3;
    float* pool = r.alloc<float>((size_t)out_rows * ldp);
    if (!pool) return TTSB_E_NOMEM;
    if (ldp != b.cin) CK(cudaMemsetAsync(pool, 0, sizeof(float) * (size_t)out_rows * ldp, r.s));
    CK(launch_pool_up2(x, ldx, pool, ldp, b.cin, b.pool_w, b.pool_b, xstats, b.n1.C, gb + b.n1.off, ld_gb, b.n1.C, 0.2f, tio.segs, nseg, tio.max_out, r.s));
    c1.x = pool; c1.ldx = ldp;
    RC(run_conv(r, c1, tout));
  }
  if (b.learned_sc) {
    float* sc = r.alloc<float>((size_t)out_rows * b.cout);
    if (!sc) return TTSB_E_NOMEM;
    ConvCall cs; cs.w = &b.sc; cs.x = x; cs.ldx = ldx; cs.y = sc; cs.ldy = b.cout; cs.tensor = tensor;
    if (b.up) { cs.in_shr = 1; RC(run_conv(r, cs, tio)); } else { RC(run_conv(r, cs, tin)); }
    scp = sc; ldsc = b.cout;
  } else if (b.up) {
    set_error("internal: upsampling block without learned shortcut"); return TTSB_E_INVALID;
  }
  ConvCall c2; c2.w = &b.conv2; c2.x = r1; c2.ldx = b.cout; c2.pad = 1; c2.y = y; c2.ldy = ldy; c2.res = scp; c2.ldr = ldsc;
  c2.pro = PRO_ADAIN_LRELU; c2.slope = 0.2f; c2.pstats = st1; c2.gb = gb; c2.ld_gb = ld_gb; c2.ad = b.n2;
  c2.out_scale = rs2; c2.ostats = ystats; c2.tensor = tensor;
  RC(run_conv(r, c2, b.up ? tout : tin));
  return TTSB_OK;
}

// kokoro AdaINResBlock1.  x (ld C) with statistics xstats; the running value lives in `cur` (ld C; may equal x when
// the caller allows in-place).  The last of the three iterations writes `(x + xt) * final_scale` into `out` (accumulating if asked).
int run_resblock1(Run& r, const ResBlock1W& b, const float* x, const double* xstats, float* cur, float* xt, float* out, float final_scale,
                  int accumulate, double* out_stats, const float* gb, int ld_gb, const SegTab& t, const SegTab* twa = nullptr, __half* cur16 = nullptr) {
  const int C = b.C, nseg = t.nseg;
  const float* in = x;
  const double* in_st = xstats;
  // twa: the 256-row tile table of the weights-as-A kernel.  On that path the block's intermediates are fp16 rows: `xt` (c1 -> c2,
  // only ever read as a 16-bit operand) and `cur16`, the running value between the three iterations (read as c1's operand and as c2's
  // residual).  The block's input and its output stay fp32, as do all statistics (taken from the values as stored).
  const bool wa = twa != nullptr;
  const SegTab& tt = wa ? *twa : t;
  if (wa && !cur16) { set_error("internal: weights-as-A resblock without an fp16 running buffer"); return TTSB_E_INVALID; }
  for (int m = 0; m < 3; ++m) {
    double* st_t = r.stats(nseg, C);
    if (!st_t) return TTSB_E_NOMEM;
    const bool in16 = wa && m > 0;                       // the running value of iterations 1, 2 lives in cur16
    ConvCall c1; c1.w = &b.c1[m]; c1.x = in16 ? reinterpret_cast<const float*>(cur16) : in; c1.ldx = C; c1.y = xt; c1.ldy = C; c1.dil = b.dil[m];
    c1.pad = (b.k * b.dil[m] - b.dil[m]) / 2;
    c1.pro = PRO_ADAIN_SNAKE; c1.pstats = in_st; c1.gb = gb; c1.ld_gb = ld_gb; c1.ad = b.a1[m]; c1.alpha = b.alpha1[m]; c1.ostats = st_t; c1.tensor = true;
    c1.wa = wa; c1.y_f16 = wa; c1.x_f16 = in16;
    RC(run_conv(r, c1, tt));
    const bool last = m == 2;
    double* st_n = last ? out_stats : r.stats(nseg, C);
    if (!last && !st_n) return TTSB_E_NOMEM;
    ConvCall c2; c2.w = &b.c2[m]; c2.x = xt; c2.ldx = C; c2.dil = 1; c2.pad = (b.k - 1) / 2;
    c2.pro = PRO_ADAIN_SNAKE; c2.pstats = st_t; c2.gb = gb; c2.ld_gb = ld_gb; c2.ad = b.a2[m]; c2.alpha = b.alpha2[m];
    c2.res = in16 ? reinterpret_cast<const float*>(cur16) : in; c2.ldr = C; c2.tensor = true;
    c2.wa = wa; c2.x_f16 = wa; c2.r_f16 = in16;
    if (last) { c2.y = out; c2.ldy = C; c2.out_scale = final_scale; c2.accumulate = accumulate; c2.ostats = out_stats; }
    else if (wa) { c2.y = reinterpret_cast<float*>(cur16); c2.ldy = C; c2.y_f16 = true; c2.ostats = st_n; }
    else { c2.y = cur; c2.ldy = C; c2.ostats = st_n; }
    RC(run_conv(r, c2, tt));
    in = cur; in_st = st_n;
  }
  return TTSB_OK;
}

// kokoro Generator.forward: gen_in [sum L, 512] + f0 [sum L] + decoder-style gamma|beta rows -> audio [300 sum L]
int run_generator(Run& r, const Space& spL, const float* gen_in, const float* f0, const float* gbD, int ld_gb, const float* noise, uint64_t seed,
                  const std::vector<unsigned long long>* utt_seeds, float* audio) {
  Engine& e = r.e;
  const int nseg = spL.n();
  std::vector<int> lU0(nseg), lS0(nseg), lU1(nseg), lS1(nseg), lA(nseg);
  for (int b = 0; b < nseg; ++b) {
    const int L = spL.len[b];
    lU0[b] = L + 1; lS0[b] = 10 * L; lU1[b] = 10 * L + 1; lS1[b] = 60 * L + 1; lA[b] = 300 * L;
  }
  Space spU0, spS0, spU1, spS1, spA;
  spU0.build(lU0); spS0.build(lS0); spU1.build(lU1); spS1.build(lS1); spA.build(lA);
  RC(r.tables_begin((size_t)nseg * 20 * sizeof(Seg) + (size_t)(spS1.total / 128 + 8 * nseg + 64) * 5 * sizeof(Tile) + (size_t)nseg * 8 + 4096));
  // STFT-frame rows of the source features: every utterance starts at a multiple of 6 rows and is zero-padded to 60L+6 rows,
  // so that six consecutive frames can be viewed as one 144-float row (noise_convs[0] without a strided kernel)
  std::vector<int> offH(nseg), lenH(nseg);
  std::vector<Seg> sgL_H(nseg), sgH_S1(nseg), sgHb_S0(nseg);
  int totH = 0;
  for (int b = 0; b < nseg; ++b) {
    const int L = spL.len[b];
    offH[b] = totH; lenH[b] = 60 * L + 1;
    sgL_H[b] = {spL.off[b], L, totH, 60 * L + 1};
    sgH_S1[b] = {totH, 60 * L + 1, spS1.off[b], 60 * L + 1};
    sgHb_S0[b] = {totH / 6, 10 * L + 1, spS0.off[b], 10 * L};
    totH += 60 * L + 6;
  }
  const SegTab tL_S1 = r.make_tab_segs(sgL_H, false);        // source + STFT
  const SegTab tH_S1 = r.make_tab_segs(sgH_S1);              // noise_convs[1] (k1)
  const SegTab tHb_S0 = r.make_tab_segs(sgHb_S0);            // noise_convs[0] over 6-frame blocks
  const SegTab tL_U0 = r.make_tab(spL, spU0);          // ups[0] polyphase
  const SegTab tU0_S0 = r.make_tab(spU0, spS0, 10, 1, false); // flat polyphase rows -> stage-0 rows
  const SegTab tS0 = r.make_tab(spS0, spS0);
  const SegTab tS0_U1 = r.make_tab(spS0, spU1);        // ups[1] polyphase
  const SegTab tU1_S1 = r.make_tab(spU1, spS1, 6, 1, false);
  const SegTab tS1 = r.make_tab(spS1, spS1);
  // stage 1 (128 channels) in the 16-bit operand modes: the weights-as-A kernel, 256-row tiles, fp16 c1 -> c2 intermediate
  static const bool no_wa = getenv("TTSB_NO_WA") != nullptr;
  const bool wa1 = !no_wa && (e.precision == TTSB_PRECISION_FP16 || e.precision == TTSB_PRECISION_BF16);
  const SegTab tS1w = wa1 ? r.make_tab(spS1, spS1, 1, 1, true, CONV_WA_TILE_ROWS) : SegTab();
  static const bool no_wa0 = getenv("TTSB_NO_WA0") != nullptr;          // A/B: stage 0 on conv_umma_kernel
  // Stage 0 is on that kernel whatever the batch: choosing by size would be faster for short inputs (one 256-row tile of 256 channels
  // is ~50 k cycles of MMAs (k11) against ~30 k for a 128-row tile of conv_umma_kernel, so with at most one tile per SM either way the
  // smaller tile finishes first: configs[0] 4.09 ms against 4.38 ms) but the two kernels keep different intermediates (fp16 against
  // fp32), and a request must give the same waveform alone and inside a batch (test_service_is_a_pure_scheduler_and_batch_invariant).
  // TTSB_NO_WA0=1 selects the 128-row kernel for a whole process (a latency-first deployment).
  const bool wa0 = wa1 && !no_wa0;
  const SegTab tS0w = wa0 ? r.make_tab(spS0, spS0, 1, 1, true, CONV_WA_TILE_ROWS) : SegTab();
  const SegTab tS1_A = r.make_tab(spS1, spA, 1, 1, false);
  // per-utterance seeds of the in-kernel SineGen noise (null: vocoder-only entry, noise keyed by position in the buffer)
  const unsigned long long* seg_seeds = (utt_seeds && (int)utt_seeds->size() == nseg) ? r.push(*utt_seeds) : nullptr;
  RC(r.tables_upload());

  // ---- harmonic source + STFT (W10, W11) ----
  float* phase = r.alloc<float>((size_t)spL.total * 9);
  float* har = r.alloc<float>((size_t)totH * 24);
  if (har) CK(cudaMemsetAsync(har, 0, sizeof(float) * (size_t)totH * 24, r.s));
  float* har_src = e.taps_on ? r.alloc<float>((size_t)spA.total) : nullptr;
  if (!phase || !har) return TTSB_E_NOMEM;
  r.e.prof.begin(r.s, PROF_SOURCE, 0.0, 0.0);
  CK(launch_source_stft(f0, noise, e.src_lin_w, e.src_lin_b, har_src, har, 24, phase, tL_S1.segs, nseg, spL.maxlen, seed, seg_seeds, r.s));
  r.e.prof.end(r.s);
  if (har_src) RC(r.tap("har_source", har_src, spA.total, 1, 1));
  RC(r.tap_ragged("har", har, 22, 24, offH, lenH));

  const float* xin = gen_in;      // stage input, ld = Cin of ups[i]
  float* stage_out = nullptr;
  for (int i = 0; i < 2; ++i) {
    const int C = i == 0 ? 256 : 128;
    const int s_up = i == 0 ? 10 : 6;
    const Space& spS = i == 0 ? spS0 : spS1;
    const Space& spU = i == 0 ? spU0 : spU1;
    const SegTab& tS = i == 0 ? tS0 : tS1;
    const size_t n = (size_t)spS.total * C;
    const size_t nup = (size_t)spU.total * s_up * C;           // the flat polyphase buffer of ups[i]: a few rows more than n
    // Four stage-size buffers serve the whole stage (liveness: xs dies with noise_res, xs2 and up die with the upsample-add):
    //   b0: xs -> up -> acc (the stage's output, kept)    b1: xt (fp16 rows on the weights-as-A path)    b2: xs2 -> cur    b3: x0
    float* b0 = r.alloc<float>(std::max(n, nup));
    const Arena::Mark stage_scope = r.mark();                   // b1..b3 (and any operand-image scratch) go back at the end of the stage
    const SegTab* twa = (i == 1 && wa1) ? &tS1w : (i == 0 && wa0) ? &tS0w : nullptr;
    float* b1 = twa ? reinterpret_cast<float*>(r.alloc<__half>(n)) : r.alloc<float>(n);
    float* b2 = r.alloc<float>(n);
    float* b3 = r.alloc<float>(n);
    double* st_xs = r.stats(nseg, C);
    if (!b0 || !b1 || !b2 || !b3 || !st_xs) return TTSB_E_NOMEM;
    // the blocks' running value between iterations (fp16 on the weights-as-A path) lives in whichever buffer is idle: x0's while
    // noise_res runs (x0 is written by the upsample-add after it), cur's while the three parallel resblocks run
    __half* c16_noise = twa ? reinterpret_cast<__half*>(b3) : nullptr;
    __half* c16_res = twa ? reinterpret_cast<__half*>(b2) : nullptr;
    // x_source = noise_res[i](noise_convs[i](har))
    float *xs = b0, *xt = b1, *xs2 = b2;
    ConvCall cn; cn.x = har; cn.y = xs; cn.ldy = C; cn.ostats = st_xs; cn.tensor = true;
    if (i == 0) { cn.w = &e.noise_conv0r; cn.ldx = 144; cn.pad = 1; RC(run_conv(r, cn, tHb_S0)); }
    else { cn.w = &e.noise_conv[1]; cn.ldx = 24; RC(run_conv(r, cn, tH_S1)); }
    RC(run_resblock1(r, e.noise_res[i], xs, st_xs, xs, xt, xs2, 1.0f, 0, nullptr, gbD, ld_gb, tS, twa, c16_noise));
    // x = ups[i](leaky_relu(x, 0.1))  (+ ReflectionPad1d((1,0)) for the last stage) + x_source
    float *up = b0, *x0 = b3;
    double* st_x0 = r.stats(nseg, C);
    if (!st_x0) return TTSB_E_NOMEM;
    ConvCall cu; cu.w = &e.ups[i]; cu.x = xin; cu.ldx = e.ups[i].Cin; cu.y = up; cu.ldy = s_up * C; cu.pad = 1; cu.pro = PRO_LRELU; cu.slope = 0.1f; cu.tensor = true;
    RC(run_conv(r, cu, i == 0 ? tL_U0 : tS0_U1));
    const SegTab& tUS = i == 0 ? tU0_S0 : tU1_S1;
    e.prof.begin(r.s, PROF_ELEMENTWISE, 0.0, 12.0 * (double)n);
    CK(launch_up_add(up, C, s_up / 2, i == 1, xs2, C, x0, C, C, tUS.segs, nseg, tUS.max_out, st_x0, C, r.s));
    e.prof.end(r.s);
    // mean of the three parallel AdaINResBlock1
    float *acc = b0, *cur = b2;
    for (int j = 0; j < 3; ++j)
      RC(run_resblock1(r, e.resblk[3 * i + j], x0, st_x0, cur, xt, acc, 1.0f / 3.0f, j > 0, nullptr, gbD, ld_gb, tS, twa, c16_res));
    r.release_to(stage_scope);
    RC(r.tap(i == 0 ? "gen_stage0" : "gen_stage1", acc, spS.total, C, C));
    xin = acc;
    stage_out = acc;
  }
  // conv_post on leaky_relu(x, 0.01), then exp / sin + iSTFT (W15, W16)
  float* cp = r.alloc<float>((size_t)spS1.total * 24);
  if (!cp) return TTSB_E_NOMEM;
  CK(cudaMemsetAsync(cp, 0, sizeof(float) * (size_t)spS1.total * 24, r.s));
  ConvCall cpc; cpc.w = &e.conv_post; cpc.x = stage_out; cpc.ldx = 128; cpc.y = cp; cpc.ldy = 24; cpc.pad = 3; cpc.pro = PRO_LRELU; cpc.slope = 0.01f; cpc.tensor = true;
  RC(run_conv(r, cpc, tS1));
  RC(r.tap("conv_post", cp, spS1.total, 22, 24));
  e.prof.begin(r.s, PROF_ISTFT, 0.0, 4.0 * (22.0 * spS1.total + (double)spA.total));
  const bool precise = e.precision == TTSB_PRECISION_FP32 || e.precision == TTSB_PRECISION_FP32_FFMA;
  CK(launch_istft(cp, 24, audio, tS1_A.segs, nseg, spS1.maxlen, precise, r.s));
  e.prof.end(r.s);
  return TTSB_OK;
}

// gamma|beta rows of every AdaIN layer of one style half: gb[b, :] = style[b, :] W^T + bias
int run_style(Run& r, const ConvW& w, const float* style, int ld_style, int nseg, float** out) {
  float* gb = r.alloc<float>((size_t)nseg * w.ldw);
  if (!gb) return TTSB_E_NOMEM;
  static const bool use_simt = getenv("TTSB_STYLE_SIMT") != nullptr;     // the tiled fp32 conv kernel: kept for bit-for-bit checks
  if (!use_simt && w.k == 1 && w.Cin <= 128) {
    r.e.prof.begin(r.s, PROF_PROJ, 2.0 * nseg * w.Cin * w.Cout, 4.0 * ((double)w.Cin * w.Cout + (double)nseg * w.Cout));
    CK(launch_style_proj(style, ld_style, nseg, w.w, w.ldw, w.bias, gb, w.ldw, w.Cin, w.Cout, r.s));
    r.e.prof.end(r.s);
    *out = gb;
    return TTSB_OK;
  }
  Space one; one.build({nseg});
  RC(r.tables_begin(4096));
  const SegTab t = r.make_tab(one, one);
  RC(r.tables_upload());
  ConvCall c; c.w = &w; c.x = style; c.ldx = ld_style; c.y = gb; c.ldy = w.ldw;
  RC(run_conv(r, c, t));
  *out = gb;
  return TTSB_OK;
}

}  // namespace

// ======================================================================================================
// begin(): token side, kmodel.py:62-119
// ======================================================================================================
int engine_begin(Engine& e, const int32_t* ids, const int32_t* token_offsets, int B, const float* ref_s, const float* speeds, int padded_lstm,
                 const int32_t* forced_dur, int32_t* pred_dur, float* duration, int32_t* frame_offsets, const int32_t* voice_slot,
                 const int32_t* voice_row) {
  if (!e.finalized) { set_error("begin: engine not finalized"); return TTSB_E_STATE; }
  const bool by_voice = voice_slot != nullptr;       // style rows come from voice packs resident on the device (ttsb_engine_voice_pack_set)
  if (B <= 0 || !ids || !token_offsets || (!ref_s && !by_voice) || (by_voice && !voice_row) || !pred_dur || !frame_offsets) {
    set_error("begin: bad arguments");
    return TTSB_E_INVALID;
  }
  if (by_voice)
    for (int b = 0; b < B; ++b) {
      const int sl = voice_slot[b];
      if (sl < 0 || sl >= (int)e.voice_pack.size() || !e.voice_pack[sl]) { set_error("begin: utterance %d names voice slot %d, which holds no pack", b, sl); return TTSB_E_INVALID; }
      if (voice_row[b] < 0 || voice_row[b] >= e.voice_pack_rows[sl]) { set_error("begin: utterance %d asks for row %d of a %d-row voice pack", b, voice_row[b], e.voice_pack_rows[sl]); return TTSB_E_INVALID; }
    }
  CK(cudaSetDevice(e.device));
  e.begun = false;
  CK(cudaStreamSynchronize(e.stream2));      // idle unless the previous step failed between a fork and its join: the workspace is about to be reused
  e.arena.reset();
  Run r(e);
  e.prof.reset();
  CK(cudaEventRecord(e.ev_step0, r.s));
  CK(cudaMemsetAsync(e.d_sat, 0, sizeof(unsigned int), r.s));
  std::vector<int> lens(B);
  int Tmax = 0;
  for (int b = 0; b < B; ++b) {
    lens[b] = token_offsets[b + 1] - token_offsets[b];
    if (lens[b] <= 0) { set_error("begin: utterance %d is empty", b); return TTSB_E_INVALID; }
    if (lens[b] > 512) { set_error("begin: utterance %d has %d tokens (> 512)", b, lens[b]); return TTSB_E_TOO_LONG; }
    Tmax = std::max(Tmax, lens[b]);
  }
  e.spT.build(lens);
  e.B = B;
  const int R = e.spT.total;
  for (int i = 0; i < R; ++i)
    if (ids[token_offsets[0] + i] < 0 || ids[token_offsets[0] + i] >= 178) { set_error("begin: token id out of range at %d", i); return TTSB_E_INVALID; }
  // a speed of 0, a negative, a denormal-small or a non-finite one would turn sum(sigmoid) / speed (kmodel.py:102) into inf or ~1e10
  // frames per token (the reference then fails inside torch and answers with silence, kmodel.py:147-150); refuse it here.
  // 50 bins / speed frames per token at most: speed >= 1e-3 bounds a token at 50 000 frames.
  if (speeds)
    for (int b = 0; b < B; ++b)
      if (!(speeds[b] >= 1e-3f) || !(speeds[b] <= 1e3f)) { set_error("begin: speed of utterance %d is %g (must be finite, in [1e-3, 1e3])", b, (double)speeds[b]); return TTSB_E_INVALID; }
  if (forced_dur)
    for (int i = 0; i < R; ++i)
      if (forced_dur[i] < 1 || forced_dur[i] > MAX_DUR_PER_TOKEN) { set_error("begin: forced_dur[%d] = %d (must be in [1, %d])", i, forced_dur[i], MAX_DUR_PER_TOKEN); return TTSB_E_INVALID; }

  // ---- tables + inputs, one upload ----
  RC(r.tables_begin((size_t)B * (sizeof(Seg) + 2 * sizeof(LstmGroup) + 260 * sizeof(float)) + (size_t)(R / 128 + B + 8) * sizeof(Tile) + (size_t)R * 16 + 8192));
  e.tabT = r.make_tab(e.spT, e.spT);
  Space spFlat; spFlat.build({R});                      // Linear layers do not care about utterance boundaries: full 128-row tiles
  const SegTab tTf = r.make_tab(spFlat, spFlat);
  std::vector<int> h_ids(ids + token_offsets[0], ids + token_offsets[0] + R), h_pos(R), h_r2s(R);
  e.utt_key.assign(B, 0);                               // FNV-1a of each utterance's token ids (engine_finish: noise seeds)
  for (int b = 0; b < B; ++b) {
    unsigned long long hsh = 0xCBF29CE484222325ull;
    for (int i = token_offsets[b]; i < token_offsets[b + 1]; ++i) { hsh ^= (unsigned long long)(unsigned)ids[i]; hsh *= 0x100000001B3ull; }
    e.utt_key[b] = hsh;
  }
  std::vector<LstmSeg> ls(B), lsq(B);
  std::vector<float> h_speed(B), h_ref;
  std::vector<const float*> h_refsrc;                    // by_voice: device address of every utterance's style row
  if (by_voice) {
    h_refsrc.resize(B);
    for (int b = 0; b < B; ++b) h_refsrc[b] = e.voice_pack[voice_slot[b]] + (size_t)voice_row[b] * 256;
  } else {
    h_ref.assign(ref_s, ref_s + (size_t)B * 256);
  }
  for (int b = 0; b < B; ++b) {
    for (int t = 0; t < lens[b]; ++t) { h_pos[e.spT.off[b] + t] = t; h_r2s[e.spT.off[b] + t] = b; }
    ls[b] = {e.spT.off[b], lens[b], 0};
    lsq[b] = {e.spT.off[b], lens[b], padded_lstm ? Tmax - lens[b] : 0};
    h_speed[b] = speeds ? speeds[b] : 1.0f;
  }
  const int* d_ids = r.push(h_ids);
  const int* d_pos = r.push(h_pos);
  const int* d_r2s = r.push(h_r2s);
  const std::vector<LstmGroup> g_ls = make_lstm_groups(ls), g_lsq = make_lstm_groups(lsq);
  const LstmGroup* d_ls = r.push(g_ls);
  const LstmGroup* d_lsq = r.push(g_lsq);
  const int n_ls = (int)g_ls.size();
  const float* d_speed = r.push(h_speed);
  const float* d_ref = by_voice ? nullptr : r.push(h_ref);
  const float* const* d_refsrc = by_voice ? r.push(h_refsrc) : nullptr;
  std::vector<int> h_forced;
  const int* d_forced = nullptr;
  if (forced_dur) { h_forced.assign(forced_dur, forced_dur + R); d_forced = r.push(h_forced); }
  RC(r.tables_upload());
  if (by_voice) {                                       // gather the rows on the device: no style bytes cross PCIe
    float* g = r.alloc<float>((size_t)B * 256);
    if (!g) return TTSB_E_NOMEM;
    CK(launch_gather_ptr_rows(d_refsrc, g, 256, B, r.s));
    d_ref = g;
  }
  e.d_refs = const_cast<float*>(d_ref);
  e.d_row2segT = const_cast<int*>(d_r2s);
  const SegTab& tT = e.tabT;

  // ---- text_encoder (A.9): independent of everything up to the length regulation - it runs on the engine's second stream beside
  // ALBERT and the duration path (whose launches fill 20-120 SMs each) and is joined before the durations are copied out ----
  static const bool one_stream_b = getenv("TTSB_ONE_STREAM") != nullptr;
  Run r2(r, one_stream_b ? r.s : e.stream2);
  if (!one_stream_b) {
    CK(cudaEventRecord(e.ev_fork, r.s));
    CK(cudaStreamWaitEvent(e.stream2, e.ev_fork, 0));
  }
  {
    float* TE0 = r.alloc<float>((size_t)R * 512);
    float* TE1 = r.alloc<float>((size_t)R * 512);
    float* TEN = r.alloc<float>((size_t)R * 512);
    float* GX2 = r.alloc<float>((size_t)R * 2048);          // its own gate scratch: the duration path's LSTMs run at the same time
    if (!TE0 || !TE1 || !TEN || !GX2) return TTSB_E_NOMEM;
    CK(launch_embedding(d_ids, e.te_emb, 512, TE0, 512, R, r2.s));
    for (int l = 0; l < 3; ++l) {
      ConvCall c; c.w = &e.te_cnn[l]; c.x = TE0; c.ldx = 512; c.y = TE1; c.ldy = 512; c.pad = 2; c.exact = true;
      RC(run_conv(r2, c, tT));
      e.prof.begin(r2.s, PROF_LAYERNORM, 0.0, 8.0 * (double)(R) * (double)(512));
      CK(launch_layernorm(TE1, 512, nullptr, 0, TE0, 512, R, 512, 1e-5f, 0, e.te_ln_g[l], e.te_ln_b[l], nullptr, 0, nullptr, ACT_LRELU, 0.2f, r2.s));
      e.prof.end(r2.s);
    }
    RC(run_bilstm(r2, e.te_lstm, TE0, 512, GX2, TEN, 512, tTf, d_ls, n_ls, true));
    e.d_ten = TEN;
    if (!one_stream_b) CK(cudaEventRecord(e.ev_join, e.stream2));
  }

  // ---- style projections of the prosody half (AdaLayerNorm + F0/N AdaIN) ----
  RC(run_style(r, e.style_pred, d_ref + 128, 256, B, &e.d_gbP));
  const int ldP = e.style_pred.ldw;

  // ---- ALBERT (A.1) ----
  float* E = r.alloc<float>((size_t)R * 128);
  float* H = r.alloc<float>((size_t)R * 768);
  float* QKV = r.alloc<float>((size_t)R * 2304);
  float* CTX = r.alloc<float>((size_t)R * 768);
  float* A1 = r.alloc<float>((size_t)R * 768);
  float* TMP = r.alloc<float>((size_t)R * 768);
  float* FF = r.alloc<float>((size_t)R * 2048);
  if (!E || !H || !QKV || !CTX || !A1 || !TMP || !FF) return TTSB_E_NOMEM;
  r.e.prof.begin(r.s, PROF_LAYERNORM, 0.0, 0.0);
  CK(launch_albert_embed(d_ids, d_pos, e.emb_word, e.emb_pos, e.emb_type, e.emb_ln_g, e.emb_ln_b, E, R, r.s));
  r.e.prof.end(r.s);
  { ConvCall c; c.w = &e.map_in; c.x = E; c.ldx = 128; c.y = H; c.ldy = 768; c.exact = true; RC(run_conv(r, c, tTf)); }
  for (int layer = 0; layer < 12; ++layer) {
    { ConvCall c; c.w = &e.qkv; c.x = H; c.ldx = 768; c.y = QKV; c.ldy = 2304; c.exact = true; RC(run_conv(r, c, tTf)); }
    r.e.prof.begin(r.s, PROF_ATTENTION, 0.0, 0.0);
    CK(launch_attention(QKV, CTX, tT.segs, B, e.spT.maxlen, 12, r.s));
    r.e.prof.end(r.s);
    { ConvCall c; c.w = &e.att_dense; c.x = CTX; c.ldx = 768; c.y = TMP; c.ldy = 768; c.res = H; c.ldr = 768; c.exact = true; RC(run_conv(r, c, tTf)); }
    r.e.prof.begin(r.s, PROF_LAYERNORM, 0.0, 8.0 * (double)(R) * (double)(768));
    CK(launch_layernorm(TMP, 768, nullptr, 0, A1, 768, R, 768, 1e-12f, 0, e.att_ln_g, e.att_ln_b, nullptr, 0, nullptr, ACT_NONE, 0.f, r.s));
    r.e.prof.end(r.s);
    { ConvCall c; c.w = &e.ffn; c.x = A1; c.ldx = 768; c.y = FF; c.ldy = 2048; c.act = ACT_GELU_NEW; c.exact = true; RC(run_conv(r, c, tTf)); }
    { ConvCall c; c.w = &e.ffn_out; c.x = FF; c.ldx = 2048; c.y = TMP; c.ldy = 768; c.res = A1; c.ldr = 768; c.exact = true; RC(run_conv(r, c, tTf)); }
    r.e.prof.begin(r.s, PROF_LAYERNORM, 0.0, 8.0 * (double)(R) * (double)(768));
    CK(launch_layernorm(TMP, 768, nullptr, 0, H, 768, R, 768, 1e-12f, 0, e.full_ln_g, e.full_ln_b, nullptr, 0, nullptr, ACT_NONE, 0.f, r.s));
    r.e.prof.end(r.s);
  }
  RC(r.tap("bert", H, R, 768, 768));

  // ---- bert_encoder + DurationEncoder (A.2, A.3) ----
  float* X = r.alloc<float>((size_t)R * 640);
  float* GX = r.alloc<float>((size_t)R * 2048);
  float* Y = r.alloc<float>((size_t)R * 512);
  if (!X || !GX || !Y) return TTSB_E_NOMEM;
  { ConvCall c; c.w = &e.bert_enc; c.x = H; c.ldx = 768; c.y = X; c.ldy = 640; c.exact = true; RC(run_conv(r, c, tTf)); }
  CK(launch_fill_style(d_ref, 256, 128, d_r2s, X, 640, 512, 128, R, r.s));
  RC(r.tap("d_en", X, R, 512, 640));
  for (int l = 0; l < 3; ++l) {
    RC(run_bilstm(r, e.de_lstm[l], X, 640, GX, Y, 512, tTf, d_ls, n_ls, true));
    r.e.prof.begin(r.s, PROF_LAYERNORM, 0.0, 8.0 * (double)(R) * (double)(512));
    CK(launch_layernorm(Y, 512, nullptr, 0, X, 640, R, 512, 1e-5f, 1, nullptr, nullptr, e.d_gbP + e.de_aln[l].off, ldP, d_r2s, ACT_NONE, 0.f, r.s));
    r.e.prof.end(r.s);
  }
  RC(r.tap("d", X, R, 640, 640));
  e.d_X = X;

  // ---- duration head (A.4) ----
  float* logits = r.alloc<float>((size_t)R * 52);
  float* d_dur = r.alloc<float>(R);
  int* d_pd = r.alloc<int>(R);
  int* d_cum = r.alloc<int>(R);
  int* d_tot = r.alloc<int>(B);
  if (!logits || !d_dur || !d_pd || !d_cum || !d_tot) return TTSB_E_NOMEM;
  RC(run_bilstm(r, e.dur_lstm, X, 640, GX, Y, 512, tTf, d_lsq, n_ls, true));
  { ConvCall c; c.w = &e.dur_proj; c.x = Y; c.ldx = 512; c.y = logits; c.ldy = 52; c.exact = true; RC(run_conv(r, c, tTf)); }
  CK(launch_duration(logits, 52, 50, d_r2s, d_speed, d_dur, d_pd, R, r.s));
  const int* d_used = d_forced ? d_forced : d_pd;
  CK(launch_dur_scan(d_used, tT.segs, B, d_cum, d_tot, r.s));
  if (!one_stream_b) CK(cudaStreamWaitEvent(r.s, e.ev_join, 0));       // the text encoder (second stream) is part of this half
  RC(r.tap("t_en", e.d_ten, R, 512, 512));
  // results the host needs to size the frame side: through PINNED memory, so that the copies are asynchronous and the text
  // encoder queued below really overlaps the host's wait (a D2H copy into pageable memory blocks the calling thread)
  const size_t out_bytes = sizeof(int) * (size_t)R + sizeof(float) * (size_t)R + sizeof(int) * (size_t)B + sizeof(unsigned int) + 64;
  if (out_bytes > e.pinned_out_cap) {
    if (e.pinned_out) cudaFreeHost(e.pinned_out);
    e.pinned_out = nullptr; e.pinned_out_cap = 0;
    const size_t cap = std::max(out_bytes * 2, size_t(1) << 16);
    CK(cudaMallocHost(&e.pinned_out, cap));
    e.pinned_out_cap = cap;
  }
  int* h_pd = reinterpret_cast<int*>(e.pinned_out);
  float* h_dur = reinterpret_cast<float*>(h_pd + R);
  int* h_tot = reinterpret_cast<int*>(h_dur + R);
  unsigned int* h_sat = reinterpret_cast<unsigned int*>(h_tot + B);
  CK(cudaMemcpyAsync(h_pd, d_pd, sizeof(int) * R, cudaMemcpyDeviceToHost, r.s));
  CK(cudaMemcpyAsync(h_dur, d_dur, sizeof(float) * R, cudaMemcpyDeviceToHost, r.s));
  CK(cudaMemcpyAsync(h_tot, d_tot, sizeof(int) * B, cudaMemcpyDeviceToHost, r.s));
  CK(cudaMemcpyAsync(h_sat, e.d_sat, sizeof(unsigned int), cudaMemcpyDeviceToHost, r.s));
  CK(cudaEventRecord(e.ev0, r.s));
  e.d_pred_dur = d_pd; e.d_cum = d_cum;

  static const bool gap_prof = getenv("TTSB_GAP_PROF") != nullptr;
  if (gap_prof) {                       // the last work of this half: the text encoder, queued to overlap the host's part
    if (!e.ev_gap0) { CK(cudaEventCreate(&e.ev_gap0)); CK(cudaEventCreate(&e.ev_gap1)); }
    CK(cudaEventRecord(e.ev_gap0, r.s));
  }

  CK(cudaEventSynchronize(e.ev0));
  memcpy(pred_dur, h_pd, sizeof(int) * R);
  if (duration) memcpy(duration, h_dur, sizeof(float) * R);
  if (*h_sat != 0) {
    set_error("begin: %u operand conversions of the split-fp16 (fp32-class) tensor-core path left fp16's range (|x| > 4094) on the text side: "
              "the durations would be wrong.  Run this checkpoint with TTSB_PRECISION_FP32_FFMA", *h_sat);
    return TTSB_E_RANGE;
  }
  std::vector<int> fl(B);
  frame_offsets[0] = 0;
  long long total = 0;
  for (int b = 0; b < B; ++b) {
    if (h_tot[b] <= 0) { set_error("begin: utterance %d has a frame count of %d (duration overflow)", b, h_tot[b]); return TTSB_E_TOO_LONG; }
    total += h_tot[b];
    if (total > MAX_FRAMES_PER_CALL) { set_error("begin: the batch needs more than %lld frames (%.1f hours of audio)", MAX_FRAMES_PER_CALL, MAX_FRAMES_PER_CALL / 40.0 / 3600.0); return TTSB_E_TOO_LONG; }
    fl[b] = h_tot[b]; frame_offsets[b + 1] = frame_offsets[b] + h_tot[b];
  }
  e.spF.build(fl);
  e.begun = true;
  return TTSB_OK;
}

// End of a step.  The strict tier (split-fp16 operands on the frame side too) always waits for the stream and reads the operand
// range counter: a saturated conversion would silently break the "fp32-class" promise, so it is an error, not a statistic.
static int finish_step(Engine& e, cudaStream_t s, bool must_sync) {
  const bool strict = e.precision == TTSB_PRECISION_FP32;
  unsigned int* h_sat = nullptr;
  if (strict) {
    if (e.pinned_out_cap < 64) {
      if (e.pinned_out) cudaFreeHost(e.pinned_out);
      e.pinned_out = nullptr; e.pinned_out_cap = 0;
      CK(cudaMallocHost(&e.pinned_out, size_t(1) << 16));
      e.pinned_out_cap = size_t(1) << 16;
    }
    h_sat = reinterpret_cast<unsigned int*>(e.pinned_out);
    CK(cudaMemcpyAsync(h_sat, e.d_sat, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
  }
  if (must_sync || strict || e.taps_on || e.prof.on) {
    CK(cudaStreamSynchronize(s));
    e.prof.collect();
  }
  if (strict && *h_sat != 0) {
    set_error("%u operand conversions of the split-fp16 (fp32-class) tensor-core path left fp16's range (|x| > 4094): the waveform is "
              "not fp32-class.  Run this checkpoint with TTSB_PRECISION_FP32_FFMA", *h_sat);
    return TTSB_E_RANGE;
  }
  return TTSB_OK;
}

// ======================================================================================================
// finish(): frame side, kmodel.py:120-143
// ======================================================================================================
int engine_finish(Engine& e, const float* noise, int noise_mem, uint64_t seed, int32_t* indices, float* audio, int audio_mem) {
  if (!e.begun) { set_error("finish: no synthesis in flight (call begin first)"); return TTSB_E_STATE; }
  if (!audio) { set_error("finish: audio is NULL"); return TTSB_E_INVALID; }
  CK(cudaSetDevice(e.device));
  e.begun = false;
  Run r(e);
  const int B = e.B;
  const Space& spT = e.spT;
  const Space& spF = e.spF;
  const int RF = spF.total;
  std::vector<int> lL(B);
  for (int b = 0; b < B; ++b) lL[b] = 2 * spF.len[b];
  Space spL; spL.build(lL);
  const int RL = spL.total;

  RC(r.tables_begin((size_t)B * 8 * sizeof(Seg) + (size_t)(RL / 128 + 4 * B + 16) * 4 * sizeof(Tile) + (size_t)RF * 4 + (size_t)B * sizeof(LstmGroup) + 8192));
  const SegTab tTF = r.make_tab(spT, spF, 1, 1, false);
  const SegTab tF = r.make_tab(spF, spF);
  const SegTab tL = r.make_tab(spL, spL);
  const SegTab tFL = r.make_tab(spF, spL);
  const SegTab tLF = r.make_tab(spL, spF, 1, 1, false);
  if (RF <= 0 || (long long)RF > MAX_FRAMES_PER_CALL) { set_error("finish: bad frame count %d", RF); return TTSB_E_INVALID; }
  std::vector<int> f2s((size_t)RF);
  std::vector<LstmSeg> lsF(B);
  for (int b = 0; b < B; ++b) {
    for (int f = 0; f < spF.len[b]; ++f) f2s[spF.off[b] + f] = b;
    lsF[b] = {spF.off[b], spF.len[b], 0};
  }
  const int* d_f2s = r.push(f2s);
  const std::vector<LstmGroup> g_lsF = make_lstm_groups(lsF);
  const LstmGroup* d_lsF = r.push(g_lsF);
  RC(r.tables_upload());
  RC(r.stats_begin(B, 40960));
  if (e.ev_gap0 && getenv("TTSB_GAP_PROF")) {
    // idle time of the GPU between the halves: from the end of begin()'s queued work to the point where finish() can launch
    CK(cudaEventRecord(e.ev_gap1, r.s));
    CK(cudaEventSynchronize(e.ev_gap1));
    float gap = 0.f;
    if (cudaEventElapsedTime(&gap, e.ev_gap0, e.ev_gap1) == cudaSuccess) fprintf(stderr, "ttsb: GPU idle between begin() and finish(): %.3f ms\n", gap);
  }

  // noise for SineGen
  const float* d_noise = nullptr;
  if (noise) {
    if (noise_mem == TTSB_MEM_DEVICE) d_noise = noise;
    else {
      float* dn = r.alloc<float>((size_t)RF * 600 * 9);
      if (!dn) return TTSB_E_NOMEM;
      CK(cudaMemcpyAsync(dn, noise, sizeof(float) * (size_t)RF * 600 * 9, cudaMemcpyHostToDevice, r.s));
      d_noise = dn;
    }
  }

  // what the generator still needs when the predictor / decoder intermediates are gone: allocated first, the rest is scoped
  float* F0 = r.alloc<float>(RL);
  float* NN = r.alloc<float>(RL);
  float* GIN = r.alloc<float>((size_t)RL * 512);
  float* gbD = nullptr;
  RC(run_style(r, e.style_dec, e.d_refs, 256, B, &gbD));
  const int ldD = e.style_dec.ldw;
  float* d_audio = audio;
  if (audio_mem != TTSB_MEM_DEVICE) d_audio = r.alloc<float>((size_t)RF * 600);
  if (!F0 || !NN || !GIN || !d_audio) return TTSB_E_NOMEM;
  const Arena::Mark front_scope = r.mark();

  // ---- alignment index map + length regulation (A.5) ----
  int* d_idx = r.alloc<int>(RF);
  float* EN = r.alloc<float>((size_t)RF * 640);
  float* CAT0 = r.alloc<float>((size_t)RF * 516);
  if (!d_idx || !EN || !CAT0) return TTSB_E_NOMEM;
  CK(launch_indices(e.d_cum, tTF.segs, B, d_idx, RF, d_f2s, r.s));
  if (indices) CK(cudaMemcpyAsync(indices, d_idx, sizeof(int) * RF, cudaMemcpyDeviceToHost, r.s));
  e.prof.begin(r.s, PROF_GATHER, 0.0, 4.0 * 640 * ((double)RF + spT.total) + 8.0 * RF);
  CK(launch_gather_rows(e.d_X, 640, d_idx, d_f2s, tTF.segs, EN, 640, 640, RF, r.s));
  e.prof.end(r.s);
  CK(cudaMemsetAsync(CAT0, 0, sizeof(float) * (size_t)RF * 516, r.s));
  e.prof.begin(r.s, PROF_GATHER, 0.0, 4.0 * 512 * ((double)RF + spT.total) + 8.0 * RF);
  CK(launch_gather_rows(e.d_ten, 512, d_idx, d_f2s, tTF.segs, CAT0, 516, 512, RF, r.s));
  e.prof.end(r.s);
  RC(r.tap("en", EN, RF, 640, 640));
  RC(r.tap("asr", CAT0, RF, 512, 516));

  // ---- F0Ntrain (A.6) ----
  const float* gbP = e.d_gbP; const int ldP = e.style_pred.ldw;
  float* GX = r.alloc<float>((size_t)RF * 2048);
  float* SH = r.alloc<float>((size_t)RF * 512);
  double* st_sh = r.stats(B, 512);
  if (!GX || !SH || !st_sh) return TTSB_E_NOMEM;
  RC(run_bilstm(r, e.shared_lstm, EN, 640, GX, SH, 512, tF, d_lsF, (int)g_lsF.size(), false));
  CK(launch_colstats(SH, 512, 512, tF.segs, B, spF.maxlen, st_sh, 512, r.s));
  // The F0 and the N branch are independent (three AdainResBlk1d + a projection each, 70-80 CTAs per launch): the N branch runs on
  // the engine's second stream, forked after the shared LSTM and joined before the decoder reads both curves.
  static const bool one_stream = getenv("TTSB_ONE_STREAM") != nullptr;
  Run r2(r, one_stream ? r.s : e.stream2);
  if (!one_stream) {
    CK(cudaEventRecord(e.ev_fork, r.s));
    CK(cudaStreamWaitEvent(e.stream2, e.ev_fork, 0));
  }
  for (int br = 0; br < 2; ++br) {
    Run& rr = br == 0 ? r : r2;
    const ResBlk1dW* blk = br == 0 ? e.f0_blk : e.n_blk;
    float* o0 = r.alloc<float>((size_t)RF * 512);
    float* o1 = r.alloc<float>((size_t)RL * 256);
    float* o2 = r.alloc<float>((size_t)RL * 256);
    double* s0 = r.stats(B, 512); double* s1 = r.stats(B, 256);
    if (!o0 || !o1 || !o2 || !s0 || !s1) return TTSB_E_NOMEM;
    RC(run_resblk1d(rr, blk[0], SH, 512, st_sh, o0, 512, s0, gbP, ldP, tF, tF, tF, true));
    RC(run_resblk1d(rr, blk[1], o0, 512, s0, o1, 256, s1, gbP, ldP, tF, tL, tFL, true));
    RC(run_resblk1d(rr, blk[2], o1, 256, s1, o2, 256, nullptr, gbP, ldP, tL, tL, tL, true));
    const ConvW& pw = br == 0 ? e.f0_proj : e.n_proj;
    static const bool proj_simt = getenv("TTSB_ROWDOT_SIMT") != nullptr;        // the tiled fp32 conv kernel: kept for bit-for-bit checks
    if (!proj_simt && pw.k == 1 && pw.Cout == 1 && pw.Cin <= 512) {
      e.prof.begin(rr.s, PROF_PROJ, 2.0 * RL * pw.Cin, 4.0 * RL * (pw.Cin + 1.0));
      CK(launch_rowdot(o2, 256, RL, pw.Cin, pw.w, pw.ldw, pw.bias, br == 0 ? F0 : NN, 1, rr.s));
      e.prof.end(rr.s);
    } else {
      ConvCall c; c.w = &pw; c.x = o2; c.ldx = 256; c.y = br == 0 ? F0 : NN; c.ldy = 1;
      RC(run_conv(rr, c, tL));
    }
  }
  if (!one_stream) {
    CK(cudaEventRecord(e.ev_join, e.stream2));
    CK(cudaStreamWaitEvent(r.s, e.ev_join, 0));
  }
  RC(r.tap("F0", F0, RL, 1, 1));
  RC(r.tap("N", NN, RL, 1, 1));

  // ---- decoder (A.10) ----
  CK(launch_conv1_s2(F0, CAT0 + 512, 516, e.f0conv_w, e.f0conv_b, tLF.segs, B, spF.maxlen, nullptr, 0, r.s));
  CK(launch_conv1_s2(NN, CAT0 + 513, 516, e.nconv_w, e.nconv_b, tLF.segs, B, spF.maxlen, nullptr, 0, r.s));
  double* st_cat0 = r.stats(B, 514);
  float* XDa = r.alloc<float>((size_t)RF * 1092);
  float* XDb = r.alloc<float>((size_t)RF * 1092);
  double* st_a = r.stats(B, 1090); double* st_b = r.stats(B, 1090);
  if (!st_cat0 || !XDa || !XDb || !st_a || !st_b) return TTSB_E_NOMEM;
  CK(launch_colstats(CAT0, 516, 514, tF.segs, B, spF.maxlen, st_cat0, 514, r.s));
  CK(cudaMemsetAsync(XDa, 0, sizeof(float) * (size_t)RF * 1092, r.s));
  CK(cudaMemsetAsync(XDb, 0, sizeof(float) * (size_t)RF * 1092, r.s));
  // tail columns [asr_res(64) | F0 | N] are the same for every decode block: write them into both ping-pong buffers
  for (float* xd : {XDa, XDb}) {
    ConvCall c; c.w = &e.asr_res; c.x = CAT0; c.ldx = 516; c.y = xd + 1024; c.ldy = 1092; c.tensor = true;
    RC(run_conv(r, c, tF));
    CK(launch_copy_cols(CAT0 + 512, 516, xd + 1088, 1092, 2, RF, r.s));
  }
  // statistics rows are [B][1090][2]: per-utterance tail statistics via one strided column-statistics pass per buffer
  CK(launch_colstats(XDa + 1024, 1092, 66, tF.segs, B, spF.maxlen, st_a + 1024 * 2, 1090, r.s));
  CK(launch_colstats(XDb + 1024, 1092, 66, tF.segs, B, spF.maxlen, st_b + 1024 * 2, 1090, r.s));
  // the conv epilogue accumulates [B][Cout=1024][2]; the AdaIN prologue of the next block reads [B][1090][2]: give the
  // producer a dense scratch and scatter it (tiny copies) — done by colstats-free strided memcpy below
  double* st_tmp = r.stats(B, 1024);
  if (!st_tmp) return TTSB_E_NOMEM;
  auto scatter_stats = [&](double* dst) -> int {
    CK(cudaMemcpy2DAsync(dst, sizeof(double) * 1090 * 2, st_tmp, sizeof(double) * 1024 * 2, sizeof(double) * 1024 * 2, B, cudaMemcpyDeviceToDevice, r.s));
    CK(cudaMemsetAsync(st_tmp, 0, sizeof(double) * (size_t)B * 1024 * 2, r.s));
    return TTSB_OK;
  };
  RC(run_resblk1d(r, e.enc_blk, CAT0, 516, st_cat0, XDa, 1092, st_tmp, gbD, ldD, tF, tF, tF, true));
  RC(scatter_stats(st_a));
  float* cur = XDa; double* cur_st = st_a; float* nxt = XDb; double* nxt_st = st_b;
  for (int i = 0; i < 3; ++i) {
    RC(run_resblk1d(r, e.dec_blk[i], cur, 1092, cur_st, nxt, 1092, st_tmp, gbD, ldD, tF, tF, tF, true));
    RC(scatter_stats(nxt_st));
    std::swap(cur, nxt); std::swap(cur_st, nxt_st);
  }
  RC(run_resblk1d(r, e.dec_blk[3], cur, 1092, cur_st, GIN, 512, nullptr, gbD, ldD, tF, tL, tFL, true));
  RC(r.tap("gen_in", GIN, RL, 512, 512));

  // ---- generator (A.11) ----
  r.release_to(front_scope);             // EN, CAT0, the LSTM gates, the F0/N and decoder intermediates: their launches are queued
  // the noise an utterance draws depends on (seed, its own token ids), not on where it sits in the batch or on its neighbours
  std::vector<unsigned long long> utt_seeds(e.utt_key.size());
  for (size_t b = 0; b < utt_seeds.size(); ++b) {
    unsigned long long z = (unsigned long long)seed + 0x9E3779B97F4A7C15ull * (e.utt_key[b] | 1ull);      // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; utt_seeds[b] = z ^ (z >> 31);
  }
  RC(run_generator(r, spL, GIN, F0, gbD, ldD, d_noise, seed, (int)utt_seeds.size() == spL.n() ? &utt_seeds : nullptr, d_audio));
  if (audio_mem != TTSB_MEM_DEVICE) CK(cudaMemcpyAsync(audio, d_audio, sizeof(float) * (size_t)RF * 600, cudaMemcpyDeviceToHost, r.s));
  CK(cudaEventRecord(e.ev_step1, r.s));
  RC(finish_step(e, r.s, audio_mem != TTSB_MEM_DEVICE || indices != nullptr));
  return TTSB_OK;
}

// ======================================================================================================
// vocoder only (BASELINE config 4)
// ======================================================================================================
int engine_vocoder(Engine& e, const float* x, const float* f0, const float* style, const int32_t* l_offsets, int B, int in_mem, const float* noise,
                   uint64_t seed, float* audio, int audio_mem) {
  if (!e.finalized) { set_error("vocoder: engine not finalized"); return TTSB_E_STATE; }
  if (B <= 0 || !x || !f0 || !style || !l_offsets || !audio) { set_error("vocoder: bad arguments"); return TTSB_E_INVALID; }
  CK(cudaSetDevice(e.device));
  e.begun = false;
  e.arena.reset();
  Run r(e);
  e.prof.reset();
  CK(cudaEventRecord(e.ev_step0, r.s));
  CK(cudaMemsetAsync(e.d_sat, 0, sizeof(unsigned int), r.s));
  std::vector<int> lL(B);
  for (int b = 0; b < B; ++b) {
    lL[b] = l_offsets[b + 1] - l_offsets[b];
    if (lL[b] <= 0) { set_error("vocoder: utterance %d is empty", b); return TTSB_E_INVALID; }
  }
  Space spL; spL.build(lL);
  const size_t RL = spL.total;
  const float *dx = x, *df0 = f0, *dstyle = style, *dnoise = noise;
  if (in_mem != TTSB_MEM_DEVICE) {
    float* a = r.alloc<float>(RL * 512); float* b = r.alloc<float>(RL); float* c = r.alloc<float>((size_t)B * 128);
    if (!a || !b || !c) return TTSB_E_NOMEM;
    CK(cudaMemcpyAsync(a, x, sizeof(float) * RL * 512, cudaMemcpyHostToDevice, r.s));
    CK(cudaMemcpyAsync(b, f0, sizeof(float) * RL, cudaMemcpyHostToDevice, r.s));
    CK(cudaMemcpyAsync(c, style, sizeof(float) * (size_t)B * 128, cudaMemcpyHostToDevice, r.s));
    dx = a; df0 = b; dstyle = c;
    if (noise) {
      float* d = r.alloc<float>(RL * 300 * 9);
      if (!d) return TTSB_E_NOMEM;
      CK(cudaMemcpyAsync(d, noise, sizeof(float) * RL * 300 * 9, cudaMemcpyHostToDevice, r.s));
      dnoise = d;
    }
  }
  float* gbD = nullptr;
  RC(run_style(r, e.style_dec, dstyle, 128, B, &gbD));
  RC(r.stats_begin(B, 16384));
  float* d_audio = audio;
  if (audio_mem != TTSB_MEM_DEVICE) {
    d_audio = r.alloc<float>(RL * 300);
    if (!d_audio) return TTSB_E_NOMEM;
  }
  RC(run_generator(r, spL, dx, df0, gbD, e.style_dec.ldw, dnoise, seed, nullptr, d_audio));   // no token ids here: noise keyed by buffer position
  if (audio_mem != TTSB_MEM_DEVICE) CK(cudaMemcpyAsync(audio, d_audio, sizeof(float) * RL * 300, cudaMemcpyDeviceToHost, r.s));
  CK(cudaEventRecord(e.ev_step1, r.s));
  RC(finish_step(e, r.s, audio_mem != TTSB_MEM_DEVICE));
  return TTSB_OK;
}

}  // namespace ttsb
