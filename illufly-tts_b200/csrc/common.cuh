// Shared declarations for the illufly-tts B200 synthesis engine (sm_100a only).
// Layout convention everywhere: activations are channels-last, rows = time steps of ALL utterances of the
// batch concatenated ("ragged rows"); a segment table says which rows belong to which utterance.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace ttsb {

// One utterance's row span on the input side and on the output side of an op.
struct Seg {
  int in_off, in_len;    // first input row, number of input rows
  int out_off, out_len;  // first output row, number of output rows
};
// One CTA's work item: segment index and first local output row.
struct Tile {
  int seg, t0;
};

enum Prologue : int {
  PRO_NONE = 0,
  PRO_LRELU = 1,         // leaky_relu(x, slope)
  PRO_ADAIN_LRELU = 2,   // leaky_relu((1+g)*(x-mean)*rstd + b, slope)
  PRO_ADAIN_SNAKE = 3,   // v=(1+g)*(x-mean)*rstd+b ; v + (1/alpha) sin^2(alpha v)
};
enum Act : int { ACT_NONE = 0, ACT_GELU_NEW = 1, ACT_LRELU = 2 };

// Implicit-GEMM 1-D convolution over ragged rows (also every Linear: k=1).
//   y[seg,t,co] = epi( bias[co] + sum_{j<k} sum_{ci<Cin} pro(x[seg, (t*stride + j*dil - pad) >> in_shr, ci]) * w[j][ci][co] )
// rows outside [0,in_len) contribute 0 (zero padding AFTER the prologue activation, as torch pads the activated tensor).
//   epi(v) = act(v + res[seg,t,co]) * out_scale (+ y_old if accumulate); optional per-(seg,co) sum / sum-of-squares
// of the stored value accumulated into ostats (double) for the next AdaIN.
struct ConvArgs {
  const float* x; int ldx;
  const float* w; int ldw;             // [k][Cin][ldw], ldw = Cout rounded up to 4
  const uint16_t* wb;                  // tensor-core path: 16-bit operand image of the weights (conv_umma.cu), else null
  int wfmt;                            // operand format of wb and of the activation slab: 0 = bf16, 1 = fp16,
                                       // 2 = split fp16 (x = hi + lo, 3 MMAs per K step: fp32-equivalent products)
  int cin_pad, cout_pad;
  const float* bias;
  float* y; int ldy;
  const float* res; int ldr;
  int Cin, Cout, k, dil, pad, stride, in_shr;
  int pro; float slope;
  const double* pstats; int ld_pstats; // [nseg][ld_pstats][2] sums of the INPUT channels (for AdaIN)
  const float* gb; int ld_gb; int beta_off;  // gamma = gb[seg*ld_gb + c], beta = gb[seg*ld_gb + beta_off + c]
  const float* alpha;                  // [Cin] Snake
  float out_scale; int accumulate; int act; float act_slope;
  float acc_scale;                     // tensor path: accumulator * acc_scale before bias (undoes the split-mode operand scaling); 0 = 1
  double* ostats; int ld_ostats;       // [nseg][ld_ostats][2]
  const Seg* segs; const Tile* tiles; int ntiles; int tile_rows;
  uint8_t* xpack; size_t xpack_bytes;  // scratch for the pre-packed split-fp16 operand images of k=1 layers (conv_umma.cu), or null
  int x_f16, y_f16, r_f16;             // conv_umma_wa.cu only: x / y / res point to fp16 rows (the intermediates of a resblock), ld* in elements
  float split_wscale;                  // wfmt 2: the power of two the weights of this layer were scaled by at pack time
  unsigned int* sat;                   // wfmt 2: device counter of operand conversions that left fp16's range (|16 x| > 65504), or null
};

struct LstmSeg {                       // host-side description of one utterance's recurrence
  int off, len, pre;                   // rows; `pre` = zero-input steps the REVERSE direction runs first (padded-batch quirk)
};

// a group of utterances whose recurrences one cluster runs together (lstm_cluster.cu); len 0 = empty slot
constexpr int LSTM_UB = 4;
struct LstmGroup {
  int off[LSTM_UB], len[LSTM_UB], pre[LSTM_UB];
};

#define TTSB_CHECK_LAUNCH() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return e__; } while (0)

// number of kernels launched by this library (bench.py reports it as gpu_launches)
extern long long g_launches;
inline cudaError_t launched(int n = 1) { __sync_fetch_and_add(&g_launches, (long long)n); return cudaGetLastError(); }

// ---- programmatic dependent launch ----------------------------------------------------------------------
// A step is ~300 launches on one stream, most of them tens of microseconds long: the gap between two launches and the prologue of
// a persistent kernel (TMEM allocation, barrier initialisation, the first weight stages) are a visible share of a short step.
// Kernels launched through launch_pdl() may start while the previous kernel on the stream is still draining; they must call
// pdl_wait() before they touch anything an earlier kernel wrote (it returns once the previous kernel has completed and flushed;
// completion is transitive because EVERY CTA of every such kernel executes pdl_wait() before it exits), and they call
// pdl_trigger() at their top so that the next kernel's launch is not held back either.  TTSB_NO_PDL=1 launches them the plain way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- kernel launchers (each returns the launch error) -------------------------------------------------
cudaError_t launch_conv_simt(const ConvArgs& a, cudaStream_t s);
cudaError_t launch_rowdot(const float* x, int ldx, int rows, int C, const float* w, int ldw, const float* bias, float* y, int ldy,
                          cudaStream_t s);            // y[r] = bias + sum_c x[r][c] w[c]: Conv1d(C, 1, 1) over flat rows
cudaError_t launch_style_proj(const float* style, int ld_style, int nseg, const float* w, int ldw, const float* bias, float* out, int ldo, int K,
                              int N, cudaStream_t s);   // out[b][n] = bias[n] + sum_k style[b][k] w[k][n], a few rows b
cudaError_t launch_conv_umma(const ConvArgs& a, cudaStream_t s);     // tcgen05 bf16 path (k*dil halo, Cin%16==0 after padding)
bool conv_umma_supported(const ConvArgs& a);
// tcgen05 path with the WEIGHTS as the A operand (M = Cout = 128 TMEM lanes, N = 256 time rows): the 128-channel layers
// (conv_umma_wa.cu).  tile_rows of the tile table must be CONV_WA_TILE_ROWS.
constexpr int CONV_WA_TILE_ROWS = 256;
cudaError_t launch_conv_umma_wa(const ConvArgs& a, cudaStream_t s);
bool conv_umma_wa_supported(const ConvArgs& a);

cudaError_t launch_lstm_cluster(const float* gx, const float* whh, const float* bias, float* y, int ldy, const LstmGroup* groups, int ngroups,
                                cudaStream_t s);
cudaError_t launch_attention(const float* qkv, float* ctx, const Seg* segs, int nseg, int max_len, int heads, cudaStream_t s);
cudaError_t launch_albert_embed(const int* ids, const int* pos, const float* word, const float* posw, const float* type0,
                                const float* g, const float* b, float* out, int rows, cudaStream_t s);
// out = LN(x (+ res)) * gamma + beta, one warp per row.  mode 0: per-channel gamma/beta (may be null);
// mode 1: AdaLayerNorm: (1+gb[seg][c]) * xhat + gb[seg][C + c], seg found through row2seg.
cudaError_t launch_layernorm(const float* x, int ldx, const float* res, int ldr, float* out, int ldo, int rows, int C, float eps,
                             int mode, const float* gamma, const float* beta, const float* gb, int ld_gb, const int* row2seg,
                             int act, float slope, cudaStream_t s);
cudaError_t launch_embedding(const int* ids, const float* table, int C, float* out, int ldo, int rows, cudaStream_t s);
cudaError_t launch_fill_style(const float* style, int ld_style, int style_off, const int* row2seg, float* out, int ldo, int col_off, int C,
                              int rows, cudaStream_t s);
cudaError_t launch_duration(const float* logits, int ld, int nbins, const int* row2seg, const float* speeds, float* duration, int* pred_dur,
                            int rows, cudaStream_t s);
cudaError_t launch_dur_scan(const int* pred_dur, const Seg* segs, int nseg, int* cum, int* totals, cudaStream_t s);
cudaError_t launch_indices(const int* cum, const Seg* segs, int nseg, int* indices, int total_frames, const int* frame2seg, cudaStream_t s);
cudaError_t launch_gather_ptr_rows(const float* const* src, float* dst, int C, int rows, cudaStream_t s);
cudaError_t launch_gather_rows(const float* src, int lds, const int* indices, const int* frame2seg, const Seg* segs, float* dst, int ldd,
                               int C, int rows, cudaStream_t s);
cudaError_t launch_colstats(const float* x, int ldx, int C, const Seg* segs, int nseg, int max_len, double* stats, int ld_stats, cudaStream_t s);
cudaError_t launch_pool_up2(const float* x, int ldx, float* y, int ldy, int C, const float* w, const float* bias,
                            const double* pstats, int ld_pstats, const float* gb, int ld_gb, int beta_off, float slope,
                            const Seg* segs, int nseg, int max_out_len, cudaStream_t s);
cudaError_t launch_conv1_s2(const float* x, float* y, int ldy, const float* w3, const float* bias1, const Seg* segs, int nseg, int max_out_len,
                            double* ostats, int ld_ostats, cudaStream_t s);
cudaError_t launch_copy_cols(const float* src, int lds, float* dst, int ldd, int C, int rows, cudaStream_t s);
cudaError_t launch_source_stft(const float* f0, const float* noise, const float* lin_w, const float* lin_b, float* har_src, float* har, int ld_har,
                               float* scratch_phase, const Seg* segs, int nseg, int max_L, unsigned long long seed, const unsigned long long* seg_seeds,
                               cudaStream_t s);
cudaError_t launch_up_add(const float* up, int ldu, int up_row_shift, int reflect, const float* xs, int ldxs, float* y, int ldy, int C,
                          const Seg* segs, int nseg, int max_out_len, double* ostats, int ld_ostats, cudaStream_t s);
cudaError_t launch_istft(const float* cp, int ld, float* audio, const Seg* segs, int nseg, int max_frames, bool precise, cudaStream_t s);
cudaError_t launch_adain_snake(const float* x, int ldx, float* y, int ldy, int C, const double* pstats, int ld_pstats, const float* gb, int ld_gb,
                               int beta_off, const float* alpha, const Seg* segs, int nseg, int max_len, cudaStream_t s);
cudaError_t launch_zero_f64(double* p, size_t n, cudaStream_t s);

// ---- device helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float snake(float v, float alpha, float inv_alpha) {
  float s = sinf(alpha * v);
  return v + inv_alpha * (s * s);
}
__device__ __forceinline__ float gelu_new(float x) {
  const float k = 0.7978845608028654f;  // sqrt(2/pi)
  float x3 = x * x * x;
  return 0.5f * x * (1.0f + tanhf(k * (x + 0.044715f * x3)));
}
__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// mean / (1+gamma)*rstd / beta from double sums — InstanceNorm1d: biased variance, eps 1e-5
__device__ __forceinline__ void adain_coef(const double* st, int n, float gamma, float beta, float& mean, float& a, float& b) {
  double m = st[0] / (double)n;
  double var = st[1] / (double)n - m * m;
  if (var < 0.0) var = 0.0;
  float rstd = (float)(1.0 / sqrt(var + 1e-5));
  mean = (float)m;
  a = (1.0f + gamma) * rstd;
  b = beta;
}

}  // namespace ttsb

#include <vector>
namespace ttsb {
// host: repack fp32 [k][Cin][ldw] weights into the bf16 shared-memory image the tcgen05 kernel streams (conv_umma.cu)
// split_scale (fmt 2 only): receives the per-layer power-of-two weight scale (ConvArgs::split_wscale)
void pack_umma_wa_weights(const float* w, int k, int C, int ldw, int fmt, std::vector<uint16_t>& out);      // conv_umma_wa.cu
void pack_umma_weights(const float* w, int k, int Cin, int Cout, int ldw, int fmt, std::vector<uint16_t>& out, int& cin_pad, int& cout_pad,
                       float* split_scale = nullptr);
const char* last_error();
}  // namespace ttsb
