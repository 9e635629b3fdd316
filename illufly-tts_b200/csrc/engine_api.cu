// C ABI (include/ttsb200.h) over the engine.  No C++ exception crosses `extern "C"`: every entry point that reaches std containers
// runs inside GUARD (std::bad_alloc -> TTSB_E_NOMEM, any other std::exception -> TTSB_E_INVALID with its message).
#include "engine.cuh"

#include <math.h>
#include <string.h>

#include <exception>
#include <new>

#include "../../include/ttsb200.h"

using namespace ttsb;

struct ttsb_engine {
  Engine e;
};

#define CKA(expr)                                                                                       \
  do {                                                                                                  \
    cudaError_t ck_e__ = (expr);                                                                        \
    if (ck_e__ != cudaSuccess) {                                                                        \
      set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(ck_e__), __FILE__, __LINE__, #expr);     \
      return TTSB_E_CUDA;                                                                               \
    }                                                                                                   \
  } while (0)

#define GUARD(where, body)                                                                               \
  try { body } catch (const std::bad_alloc&) { set_error(where ": out of host memory"); return TTSB_E_NOMEM; } \
  catch (const std::exception& ex) { set_error(where ": %s", ex.what()); return TTSB_E_INVALID; }              \
  catch (...) { set_error(where ": unknown C++ exception"); return TTSB_E_INVALID; }

#define GUARD_END(where)                                                                                 \
  catch (const std::bad_alloc&) { set_error(where ": out of host memory"); return TTSB_E_NOMEM; }        \
  catch (const std::exception& ex) { set_error(where ": %s", ex.what()); return TTSB_E_INVALID; }        \
  catch (...) { set_error(where ": unknown C++ exception"); return TTSB_E_INVALID; }

namespace {

struct OpTimer {
  Engine& e; float* ms; bool on;
  OpTimer(Engine& en, float* m) : e(en), ms(m), on(m != nullptr) {}
  cudaError_t start() { return on ? cudaEventRecord(e.ev0, e.stream) : cudaSuccess; }
  cudaError_t stop() {
    if (!on) return cudaSuccess;
    cudaError_t r;
    if ((r = cudaEventRecord(e.ev1, e.stream)) != cudaSuccess) return r;
    if ((r = cudaEventSynchronize(e.ev1)) != cudaSuccess) return r;
    return cudaEventElapsedTime(ms, e.ev0, e.ev1);
  }
};

// segment + tile tables for a same-space op, uploaded synchronously (op entry points are not latency critical)
struct OpTables {
  Seg* segs = nullptr; Tile* tiles = nullptr; int ntiles = 0, maxlen = 0, total = 0;
  int build(Engine& e, const std::vector<int>& in_len, const std::vector<int>& out_len) {
    const int B = (int)in_len.size();
    std::vector<Seg> s(B); std::vector<Tile> t;
    int io = 0, oo = 0;
    for (int b = 0; b < B; ++b) {
      s[b] = {io, in_len[b], oo, out_len[b]};
      for (int t0 = 0; t0 < out_len[b]; t0 += 128) t.push_back({b, t0});
      io += in_len[b]; oo += out_len[b];
      maxlen = std::max(maxlen, out_len[b]);
    }
    total = oo; ntiles = (int)t.size();
    segs = (Seg*)e.arena.alloc(sizeof(Seg) * B);
    tiles = (Tile*)e.arena.alloc(sizeof(Tile) * std::max<size_t>(t.size(), 1));
    if (!segs || !tiles) { set_error("workspace allocation failed"); return TTSB_E_NOMEM; }
    CKA(cudaMemcpyAsync(segs, s.data(), sizeof(Seg) * B, cudaMemcpyHostToDevice, e.stream));
    if (!t.empty()) CKA(cudaMemcpyAsync(tiles, t.data(), sizeof(Tile) * t.size(), cudaMemcpyHostToDevice, e.stream));
    CKA(cudaStreamSynchronize(e.stream));
    return TTSB_OK;
  }
};

std::vector<int> lens_from_offsets(const int32_t* off, int B) {
  std::vector<int> l(B);
  for (int b = 0; b < B; ++b) l[b] = off[b + 1] - off[b];
  return l;
}

}  // namespace

extern "C" {

int ttsb_abi_version(void) { return TTSB_ABI_VERSION; }
const char* ttsb_last_error(void) { return last_error(); }

int ttsb_engine_create(int cuda_device, ttsb_engine** out) {
  if (!out) { set_error("create: out is NULL"); return TTSB_E_INVALID; }
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    set_error("create: no CUDA device visible (this engine has no CPU fallback)");
    return TTSB_E_CUDA;
  }
  if (cuda_device < 0 || cuda_device >= n) { set_error("create: device %d out of range (%d visible)", cuda_device, n); return TTSB_E_INVALID; }
  CKA(cudaSetDevice(cuda_device));
  cudaDeviceProp prop;
  CKA(cudaGetDeviceProperties(&prop, cuda_device));
  // the library carries one sm_100a cubin and no PTX: sm_103 (B300) would fail later with "no kernel image"
  if (prop.major != 10 || prop.minor != 0) { set_error("create: device %d is sm_%d%d; this library is built for sm_100a only", cuda_device, prop.major, prop.minor); return TTSB_E_CUDA; }
  ttsb_engine* h = nullptr;
  GUARD("create", h = new ttsb_engine();)
  h->e.device = cuda_device;
  auto setup = [&]() -> int {
    CKA(cudaStreamCreateWithFlags(&h->e.stream, cudaStreamNonBlocking));
    CKA(cudaStreamCreateWithFlags(&h->e.stream2, cudaStreamNonBlocking));
    CKA(cudaEventCreateWithFlags(&h->e.ev_fork, cudaEventDisableTiming));
    CKA(cudaEventCreateWithFlags(&h->e.ev_join, cudaEventDisableTiming));
    CKA(cudaEventCreate(&h->e.ev0));
    CKA(cudaEventCreate(&h->e.ev1));
    CKA(cudaEventCreateWithFlags(&h->e.ev_sync, cudaEventDisableTiming));
    CKA(cudaEventCreate(&h->e.ev_step0));
    CKA(cudaEventCreate(&h->e.ev_step1));
    CKA(cudaMalloc(&h->e.d_sat, 256));
    CKA(cudaMemset(h->e.d_sat, 0, 256));
    return TTSB_OK;
  };
  const int rc = setup();
  if (rc != TTSB_OK) { ttsb_engine_destroy(h); return rc; }      // whatever was created so far is released (the message stays)
  *out = h;
  return TTSB_OK;
}

int ttsb_engine_destroy(ttsb_engine* h) {
  if (!h) return TTSB_OK;
  Engine& e = h->e;
  cudaSetDevice(e.device);
  if (e.stream) cudaStreamSynchronize(e.stream);
  if (e.stream2) cudaStreamSynchronize(e.stream2);
  for (void* p : e.dev_allocs) cudaFree(p);
  for (float* p : e.voice_pack) if (p) cudaFree(p);
  for (auto& kv : e.taps) if (kv.second.d) cudaFree(kv.second.d);
  e.arena.release();
  if (e.pinned) cudaFreeHost(e.pinned);
  if (e.pinned_out) cudaFreeHost(e.pinned_out);
  if (e.d_sat) cudaFree(e.d_sat);
  if (e.ev0) cudaEventDestroy(e.ev0);
  if (e.ev1) cudaEventDestroy(e.ev1);
  if (e.ev_sync) cudaEventDestroy(e.ev_sync);
  if (e.ev_step0) cudaEventDestroy(e.ev_step0);
  if (e.ev_step1) cudaEventDestroy(e.ev_step1);
  e.prof.release();
  if (e.ev_fork) cudaEventDestroy(e.ev_fork);
  if (e.ev_join) cudaEventDestroy(e.ev_join);
  if (e.stream2) cudaStreamDestroy(e.stream2);
  if (e.stream) cudaStreamDestroy(e.stream);
  delete h;
  return TTSB_OK;
}

int ttsb_engine_set_tensor(ttsb_engine* h, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!h || !name || !data || !shape || ndim < 1 || ndim > 4) { set_error("set_tensor: bad arguments"); return TTSB_E_INVALID; }
  if (h->e.finalized) { set_error("set_tensor: engine already finalized"); return TTSB_E_STATE; }
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  if (n <= 0) { set_error("set_tensor: '%s' is empty", name); return TTSB_E_INVALID; }
  GUARD("set_tensor",
    t.data.assign(data, data + n);
    h->e.host[name] = std::move(t);)
  return TTSB_OK;
}

int ttsb_engine_finalize(ttsb_engine* h) {
  if (!h) { set_error("finalize: engine is NULL"); return TTSB_E_INVALID; }
  CKA(cudaSetDevice(h->e.device));
  GUARD("finalize", return engine_finalize(h->e);)
}

int ttsb_engine_set_precision(ttsb_engine* h, int precision) {
  if (!h || precision < TTSB_PRECISION_FP32 || precision > TTSB_PRECISION_FP32_FFMA) { set_error("set_precision: bad arguments"); return TTSB_E_INVALID; }
  h->e.precision = precision;
  return TTSB_OK;
}
int ttsb_engine_get_precision(ttsb_engine* h) { return h ? h->e.precision : TTSB_E_INVALID; }
int ttsb_engine_enable_taps(ttsb_engine* h, int on) {
  if (!h) return TTSB_E_INVALID;
  h->e.taps_on = on != 0;
  return TTSB_OK;
}
int64_t ttsb_engine_launch_count(ttsb_engine* h) { (void)h; return (int64_t)g_launches; }

int ttsb_engine_workspace_peak(ttsb_engine* h, int64_t* bytes) {
  if (!h || !bytes) { set_error("workspace_peak: bad arguments"); return TTSB_E_INVALID; }
  *bytes = (int64_t)h->e.arena.peak();
  return TTSB_OK;
}

int ttsb_engine_last_step_ms(ttsb_engine* h, float* ms) {
  if (!h || !ms) { set_error("last_step_ms: bad arguments"); return TTSB_E_INVALID; }
  CKA(cudaSetDevice(h->e.device));
  CKA(cudaEventSynchronize(h->e.ev_step1));
  CKA(cudaEventElapsedTime(ms, h->e.ev_step0, h->e.ev_step1));
  return TTSB_OK;
}
int ttsb_engine_profile(ttsb_engine* h, int on) {
  if (!h) return TTSB_E_INVALID;
  h->e.prof.on = on != 0;
  return TTSB_OK;
}
int ttsb_engine_profile_read(ttsb_engine* h, int category, double* ms, double* flops, double* bytes, int64_t* launches) {
  if (!h || category < 0 || category >= PROF_NCAT) { set_error("profile_read: bad arguments"); return TTSB_E_INVALID; }
  const Profiler& p = h->e.prof;
  if (ms) *ms = p.ms[category];
  if (flops) *flops = p.flops[category];
  if (bytes) *bytes = p.bytes[category];
  if (launches) *launches = p.launches[category];
  return TTSB_OK;
}

int ttsb_synthesize_begin(ttsb_engine* h, const int32_t* ids, const int32_t* token_offsets, int B, const float* ref_s, const float* speeds,
                          int padded_lstm, const int32_t* forced_dur, int32_t* pred_dur, float* duration, int32_t* frame_offsets) {
  if (!h) { set_error("begin: engine is NULL"); return TTSB_E_INVALID; }
  GUARD("begin", return engine_begin(h->e, ids, token_offsets, B, ref_s, speeds, padded_lstm, forced_dur, pred_dur, duration, frame_offsets);)
}

// ---- voice packs resident on the device ----
constexpr int TTSB_MAX_VOICE_SLOTS = 4096;
int ttsb_engine_voice_pack_set(ttsb_engine* h, int slot, const float* pack, int rows) {
  if (!h) { set_error("voice_pack_set: engine is NULL"); return TTSB_E_INVALID; }
  GUARD("voice_pack_set",
    Engine& e = h->e;
    if (slot < 0 || slot >= TTSB_MAX_VOICE_SLOTS || (pack && (rows < 1 || rows > 65536))) { set_error("voice_pack_set: slot %d / rows %d out of range", slot, rows); return TTSB_E_INVALID; }
    CKA(cudaSetDevice(e.device));
    // a step in flight has gathered its rows before begin() returned (begin ends with a host synchronisation), so a pack can be replaced
    // between calls; callers of one engine are serialised by the host side
    if ((int)e.voice_pack.size() <= slot) { e.voice_pack.resize(slot + 1, nullptr); e.voice_pack_rows.resize(slot + 1, 0); }
    if (e.voice_pack[slot]) { CKA(cudaStreamSynchronize(e.stream)); cudaFree(e.voice_pack[slot]); e.voice_pack[slot] = nullptr; e.voice_pack_rows[slot] = 0; }
    if (!pack) return TTSB_OK;                           // NULL pack: the slot is emptied
    float* d = nullptr;
    if (cudaMalloc(&d, sizeof(float) * (size_t)rows * 256) != cudaSuccess) { cudaGetLastError(); set_error("voice_pack_set: out of device memory"); return TTSB_E_NOMEM; }
    if (cudaMemcpy(d, pack, sizeof(float) * (size_t)rows * 256, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); cudaGetLastError(); set_error("voice_pack_set: copy failed"); return TTSB_E_CUDA; }
    e.voice_pack[slot] = d; e.voice_pack_rows[slot] = rows;
    return TTSB_OK;)
}

int ttsb_synthesize_begin_voices(ttsb_engine* h, const int32_t* ids, const int32_t* token_offsets, int B, const int32_t* voice_slot, const int32_t* voice_row,
                                 const float* speeds, int padded_lstm, const int32_t* forced_dur, int32_t* pred_dur, float* duration, int32_t* frame_offsets) {
  if (!h) { set_error("begin_voices: engine is NULL"); return TTSB_E_INVALID; }
  if (!voice_slot || !voice_row) { set_error("begin_voices: voice_slot / voice_row is NULL"); return TTSB_E_INVALID; }
  GUARD("begin_voices", return engine_begin(h->e, ids, token_offsets, B, nullptr, speeds, padded_lstm, forced_dur, pred_dur, duration, frame_offsets, voice_slot, voice_row);)
}

int ttsb_synthesize_finish(ttsb_engine* h, const float* noise, int noise_mem, uint64_t seed, int32_t* indices, float* audio, int audio_mem) {
  if (!h) { set_error("finish: engine is NULL"); return TTSB_E_INVALID; }
  GUARD("finish", return engine_finish(h->e, noise, noise_mem, seed, indices, audio, audio_mem);)
}

int ttsb_vocoder(ttsb_engine* h, const float* x, const float* f0, const float* style, const int32_t* l_offsets, int B, int in_mem, const float* noise,
                 uint64_t seed, float* audio, int audio_mem) {
  if (!h) { set_error("vocoder: engine is NULL"); return TTSB_E_INVALID; }
  GUARD("vocoder", return engine_vocoder(h->e, x, f0, style, l_offsets, B, in_mem, noise, seed, audio, audio_mem);)
}

int ttsb_engine_read_tap(ttsb_engine* h, const char* name, float* out, int64_t* rows, int64_t* cols) {
  if (!h || !name) { set_error("read_tap: bad arguments"); return TTSB_E_INVALID; }
  auto it = h->e.taps.find(name);
  if (it == h->e.taps.end()) { set_error("read_tap: no tap named '%s'", name); return TTSB_E_INVALID; }
  if (rows) *rows = it->second.rows;
  if (cols) *cols = it->second.cols;
  if (out && it->second.d) {
    CKA(cudaSetDevice(h->e.device));
    CKA(cudaStreamSynchronize(h->e.stream));
    CKA(cudaMemcpy(out, it->second.d, sizeof(float) * it->second.rows * it->second.cols, cudaMemcpyDeviceToHost));
  }
  return TTSB_OK;
}

// ---------------------------------------------------------------------------------------------------------
// single-kernel entry points
// ---------------------------------------------------------------------------------------------------------
int ttsb_op_istft(ttsb_engine* h, const float* cp, int ld, const int32_t* frame_counts_F, int B, float* audio, float* ms) try {
  if (!h || !cp || !frame_counts_F || B <= 0 || !audio || ld < 24 || (ld & 3)) { set_error("op_istft: bad arguments (rows must hold >= 24 floats, 16-byte aligned)"); return TTSB_E_INVALID; }
  Engine& e = h->e;
  CKA(cudaSetDevice(e.device));
  e.begun = false; e.arena.reset();
  std::vector<int> fi(B), fo(B);
  for (int b = 0; b < B; ++b) { fi[b] = 120 * frame_counts_F[b] + 1; fo[b] = 600 * frame_counts_F[b]; }
  OpTables t;
  int rc = t.build(e, fi, fo);
  if (rc) return rc;
  int maxf = 0;
  for (int b = 0; b < B; ++b) maxf = std::max(maxf, fi[b]);
  OpTimer tm(e, ms);
  CKA(tm.start());
  CKA(launch_istft(cp, ld, audio, t.segs, B, maxf, e.precision == TTSB_PRECISION_FP32 || e.precision == TTSB_PRECISION_FP32_FFMA, e.stream));
  CKA(tm.stop());
  return TTSB_OK;
} GUARD_END("op_istft")

int ttsb_op_adain_snake(ttsb_engine* h, const float* x, float* y, int C, const int32_t* row_offsets, int B, const float* gamma_beta, const float* alpha,
                        float* ms_stats, float* ms_apply) try {
  if (!h || !x || !y || !row_offsets || B <= 0 || !gamma_beta || !alpha || C <= 0 || (C & 3)) { set_error("op_adain_snake: bad arguments"); return TTSB_E_INVALID; }
  Engine& e = h->e;
  CKA(cudaSetDevice(e.device));
  e.begun = false; e.arena.reset();
  const std::vector<int> l = lens_from_offsets(row_offsets, B);
  OpTables t;
  int rc = t.build(e, l, l);
  if (rc) return rc;
  double* st = (double*)e.arena.alloc(sizeof(double) * (size_t)B * C * 2);
  if (!st) { set_error("workspace allocation failed"); return TTSB_E_NOMEM; }
  CKA(cudaMemsetAsync(st, 0, sizeof(double) * (size_t)B * C * 2, e.stream));
  {
    OpTimer tm(e, ms_stats);
    CKA(tm.start());
    CKA(launch_colstats(x, C, C, t.segs, B, t.maxlen, st, C, e.stream));
    CKA(tm.stop());
  }
  {
    OpTimer tm(e, ms_apply);
    CKA(tm.start());
    CKA(launch_adain_snake(x, C, y, C, C, st, C, gamma_beta, 2 * C, C, alpha, t.segs, B, t.maxlen, e.stream));
    CKA(tm.stop());
  }
  CKA(cudaStreamSynchronize(e.stream));
  return TTSB_OK;
} GUARD_END("op_adain_snake")

int ttsb_op_length_regulate(ttsb_engine* h, const int32_t* pred_dur, const int32_t* token_offsets, int B, const float* src, int C, int32_t* indices,
                            float* dst, int32_t* frame_offsets_host, float* ms) try {
  if (!h || !pred_dur || !token_offsets || B <= 0 || !src || !indices || !dst || !frame_offsets_host || C <= 0 || (C & 3)) {
    set_error("op_length_regulate: bad arguments"); return TTSB_E_INVALID;
  }
  Engine& e = h->e;
  CKA(cudaSetDevice(e.device));
  e.begun = false; e.arena.reset();
  const std::vector<int> lt = lens_from_offsets(token_offsets, B);
  OpTables tt;
  int rc = tt.build(e, lt, lt);
  if (rc) return rc;
  const int R = tt.total;
  int* cum = (int*)e.arena.alloc(sizeof(int) * R);
  int* tot = (int*)e.arena.alloc(sizeof(int) * B);
  if (!cum || !tot) { set_error("workspace allocation failed"); return TTSB_E_NOMEM; }
  CKA(launch_dur_scan(pred_dur, tt.segs, B, cum, tot, e.stream));
  std::vector<int> h_tot(B);
  CKA(cudaMemcpyAsync(h_tot.data(), tot, sizeof(int) * B, cudaMemcpyDeviceToHost, e.stream));
  CKA(cudaStreamSynchronize(e.stream));
  OpTables tf;
  rc = tf.build(e, lt, h_tot);
  if (rc) return rc;
  const int RF = tf.total;
  frame_offsets_host[0] = 0;
  std::vector<int> f2s(RF);
  for (int b = 0, f = 0; b < B; ++b) {
    frame_offsets_host[b + 1] = frame_offsets_host[b] + h_tot[b];
    for (int i = 0; i < h_tot[b]; ++i) f2s[f++] = b;
  }
  int* d_f2s = (int*)e.arena.alloc(sizeof(int) * std::max(RF, 1));
  if (!d_f2s) { set_error("workspace allocation failed"); return TTSB_E_NOMEM; }
  CKA(cudaMemcpyAsync(d_f2s, f2s.data(), sizeof(int) * RF, cudaMemcpyHostToDevice, e.stream));
  CKA(cudaStreamSynchronize(e.stream));
  OpTimer tm(e, ms);
  CKA(tm.start());
  CKA(launch_indices(cum, tf.segs, B, indices, RF, d_f2s, e.stream));
  CKA(launch_gather_rows(src, C, indices, d_f2s, tf.segs, dst, C, C, RF, e.stream));
  CKA(tm.stop());
  CKA(cudaStreamSynchronize(e.stream));
  return TTSB_OK;
} GUARD_END("op_length_regulate")

int ttsb_op_conv1d(ttsb_engine* h, const float* x, const float* w, const float* bias, float* y, int Cin, int Cout, int k, int dilation,
                   const int32_t* row_offsets, int B, int precision, int reps, float* ms) try {
  if (!h || !x || !w || !y || !row_offsets || B <= 0 || Cin <= 0 || Cout <= 0 || k <= 0 || dilation <= 0 || (Cin & 3)) {
    set_error("op_conv1d: bad arguments"); return TTSB_E_INVALID;
  }
  Engine& e = h->e;
  CKA(cudaSetDevice(e.device));
  e.begun = false; e.arena.reset();
  const std::vector<int> l = lens_from_offsets(row_offsets, B);
  OpTables t;
  int rc = t.build(e, l, l);
  if (rc) return rc;
  // repack [Cout,Cin,k] -> [k][Cin][ldw] (+ bf16 image), upload into the workspace
  const int ldw = (Cout + 3) & ~3;
  std::vector<float> p((size_t)k * Cin * ldw, 0.f);
  for (int co = 0; co < Cout; ++co)
    for (int ci = 0; ci < Cin; ++ci)
      for (int j = 0; j < k; ++j) p[((size_t)j * Cin + ci) * ldw + co] = w[((size_t)co * Cin + ci) * k + j];
  std::vector<uint16_t> pb;
  int cin_pad = 0, cout_pad = 0;
  // as the engine: FP32 = tcgen05 with split-fp16 operands (the FFMA kernel where the shape is unsupported), FP32_FFMA = CUDA cores
  const int wfmt = precision == TTSB_PRECISION_FP32 ? 2 : (precision == TTSB_PRECISION_FP16 ? 1 : 0);
  float wscale = 0.f;
  pack_umma_weights(p.data(), k, Cin, Cout, ldw, wfmt, pb, cin_pad, cout_pad, &wscale);
  float* dw = (float*)e.arena.alloc(sizeof(float) * p.size());
  uint16_t* dwb = (uint16_t*)e.arena.alloc(sizeof(uint16_t) * std::max<size_t>(pb.size(), 8));
  float* db = bias ? (float*)e.arena.alloc(sizeof(float) * Cout) : nullptr;
  if (!dw || !dwb || (bias && !db)) { set_error("workspace allocation failed"); return TTSB_E_NOMEM; }
  CKA(cudaMemcpyAsync(dw, p.data(), sizeof(float) * p.size(), cudaMemcpyHostToDevice, e.stream));
  if (!pb.empty()) CKA(cudaMemcpyAsync(dwb, pb.data(), sizeof(uint16_t) * pb.size(), cudaMemcpyHostToDevice, e.stream));
  if (bias) CKA(cudaMemcpyAsync(db, bias, sizeof(float) * Cout, cudaMemcpyHostToDevice, e.stream));
  CKA(cudaStreamSynchronize(e.stream));
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.ldx = Cin; a.w = dw; a.ldw = ldw; a.wb = pb.empty() ? nullptr : dwb; a.wfmt = wfmt; a.cin_pad = cin_pad; a.cout_pad = cout_pad; a.bias = db;
  a.y = y; a.ldy = Cout; a.Cin = Cin; a.Cout = Cout; a.k = k; a.dil = dilation; a.pad = (k * dilation - dilation) / 2; a.stride = 1;
  a.out_scale = 1.f; a.segs = t.segs; a.tiles = t.tiles; a.ntiles = t.ntiles; a.tile_rows = 128;
  a.split_wscale = wscale; a.sat = wfmt == 2 ? e.d_sat : nullptr;
  bool tensor = precision != TTSB_PRECISION_FP32_FFMA;
  if (tensor && !conv_umma_supported(a)) {
    if (precision != TTSB_PRECISION_FP32) { set_error("op_conv1d: shape not supported by the tcgen05 kernel"); return TTSB_E_INVALID; }
    tensor = false;
  }
  if (!tensor) { a.wb = nullptr; a.wfmt = 0; a.sat = nullptr; }
  if (reps < 1) reps = 1;
  if (a.sat) CKA(cudaMemsetAsync(a.sat, 0, sizeof(unsigned int), e.stream));
  OpTimer tm(e, ms);
  CKA(tm.start());
  for (int i = 0; i < reps; ++i) {
    if (tensor) CKA(launch_conv_umma(a, e.stream)); else CKA(launch_conv_simt(a, e.stream));
  }
  CKA(tm.stop());
  if (ms) *ms /= (float)reps;
  CKA(cudaStreamSynchronize(e.stream));
  if (a.sat) {      // the split-fp16 path's range check, as the engine does after every step of the strict tier
    unsigned int n_sat = 0;
    CKA(cudaMemcpy(&n_sat, a.sat, sizeof(n_sat), cudaMemcpyDeviceToHost));
    if (n_sat) { set_error("op_conv1d: %u operand conversions left fp16's range (|x| > 4094) in the split-fp16 path", n_sat); return TTSB_E_RANGE; }
  }
  return TTSB_OK;
} GUARD_END("op_conv1d")

int ttsb_op_bilstm(ttsb_engine* h, const float* x, int In, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                   const int32_t* row_offsets, int B, float* y, float* ms) try {
  if (!h || !x || !w_ih || !w_hh || !b_ih || !b_hh || !row_offsets || B <= 0 || !y || In <= 0 || (In & 3)) { set_error("op_bilstm: bad arguments"); return TTSB_E_INVALID; }
  Engine& e = h->e;
  CKA(cudaSetDevice(e.device));
  e.begun = false; e.arena.reset();
  const std::vector<int> l = lens_from_offsets(row_offsets, B);
  OpTables t;
  int rc = t.build(e, l, l);
  if (rc) return rc;
  std::vector<float> p((size_t)In * 2048), bias(2048);
  for (int g = 0; g < 2048; ++g) {
    bias[g] = b_ih[g] + b_hh[g];
    for (int ci = 0; ci < In; ++ci) p[(size_t)ci * 2048 + g] = w_ih[(size_t)g * In + ci];
  }
  std::vector<LstmGroup> ls((B + LSTM_UB - 1) / LSTM_UB);
  for (auto& g : ls) memset(&g, 0, sizeof(g));
  for (int b = 0, o = 0; b < B; ++b) { ls[b / LSTM_UB].off[b % LSTM_UB] = o; ls[b / LSTM_UB].len[b % LSTM_UB] = l[b]; o += l[b]; }
  std::vector<float> whh_o(w_hh, w_hh + (size_t)2 * 1024 * 256);
  float* dw = (float*)e.arena.alloc(sizeof(float) * p.size());
  float* db = (float*)e.arena.alloc(sizeof(float) * 2048);
  float* dh = (float*)e.arena.alloc(sizeof(float) * whh_o.size());
  float* gx = (float*)e.arena.alloc(sizeof(float) * (size_t)t.total * 2048);
  LstmGroup* dls = (LstmGroup*)e.arena.alloc(sizeof(LstmGroup) * ls.size());
  if (!dw || !db || !dh || !gx || !dls) { set_error("workspace allocation failed"); return TTSB_E_NOMEM; }
  CKA(cudaMemcpyAsync(dw, p.data(), sizeof(float) * p.size(), cudaMemcpyHostToDevice, e.stream));
  CKA(cudaMemcpyAsync(db, bias.data(), sizeof(float) * 2048, cudaMemcpyHostToDevice, e.stream));
  CKA(cudaMemcpyAsync(dh, whh_o.data(), sizeof(float) * whh_o.size(), cudaMemcpyHostToDevice, e.stream));
  CKA(cudaMemcpyAsync(dls, ls.data(), sizeof(LstmGroup) * ls.size(), cudaMemcpyHostToDevice, e.stream));
  CKA(cudaStreamSynchronize(e.stream));
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.ldx = In; a.w = dw; a.ldw = 2048; a.bias = db; a.y = gx; a.ldy = 2048; a.Cin = In; a.Cout = 2048; a.k = 1; a.dil = 1; a.stride = 1;
  a.out_scale = 1.f; a.segs = t.segs; a.tiles = t.tiles; a.ntiles = t.ntiles; a.tile_rows = 128;
  OpTimer tm(e, ms);
  CKA(tm.start());
  CKA(launch_conv_simt(a, e.stream));
  CKA(launch_lstm_cluster(gx, dh, db, y, 512, dls, (int)ls.size(), e.stream));
  CKA(tm.stop());
  CKA(cudaStreamSynchronize(e.stream));
  return TTSB_OK;
} GUARD_END("op_bilstm")

}  // extern "C"
