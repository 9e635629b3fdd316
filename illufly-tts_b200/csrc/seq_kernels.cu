// Text-side kernels: ALBERT embedding + LayerNorms, attention, BiLSTM recurrence, duration head,
// duration scan / alignment index map / length-regulation gather.  All fp32: this side decides the integer
// durations, which must match the reference bit-for-bit.
#include "common.cuh"

namespace ttsb {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm family: one warp per row, row held in registers (C <= 32*MAXV), two-pass variance.
// ---------------------------------------------------------------------------------------------------------
constexpr int MAXV = 24;  // 768 / 32

__global__ void __launch_bounds__(256, 2) layernorm_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ res, int ldr,
                                                        float* __restrict__ out, int ldo, int rows, int C, float eps, int mode,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ gb, int ld_gb, const int* __restrict__ row2seg,
                                                        int act, float slope) {
  pdl_trigger();
  pdl_wait();               // before any return: every CTA waits for the previous kernel (common.cuh, programmatic dependent launch)
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // Every load of a pass is issued before anything waits on one.  A warp issues in order, and a load behind `if (c < C)` (C is a
  // run-time value) cannot be hoisted above its branch: with the loads inside the element loop, the 24 loads of a lane were 24
  // SERIAL round trips (ncu: 88 % of the stalls on the long scoreboard at two lines, 10 % of the issue slots busy, 22 us for a
  // 15-MB LayerNorm).  The addresses are clamped into the row instead and the value is selected afterwards; the arithmetic and its
  // order are unchanged, so the results are bit-identical.
  const int seg = (mode == 1) ? row2seg[row] : 0;
  const float* gp = (mode == 1) ? gb + (size_t)seg * ld_gb : gamma;          // scale row (AdaLayerNorm: gamma | beta of the utterance)
  const float* bp = (mode == 1) ? gb + (size_t)seg * ld_gb + C : beta;
  const bool affine = (mode == 1) || gamma != nullptr;
  float v[MAXV], gv[MAXV], bv[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    const int cc = c < C ? c : C - 1;
    v[i] = x[(size_t)row * ldx + cc];
  }
  if (res) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      const int cc = c < C ? c : C - 1;
      gv[i] = res[(size_t)row * ldr + cc];                 // borrowed as a temporary
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) v[i] += gv[i];
  }
  if (affine) {                                            // in flight during the two reductions below
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = lane + i * 32;
      const int cc = c < C ? c : C - 1;
      gv[i] = gp[cc];
      bv[i] = bp[cc];
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    const float t = c < C ? v[i] : 0.f;
    v[i] = t;
    s += t;
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < C) { const float d = v[i] - mean; q += d * d; }
  }
  const float var = warp_sum(q) / (float)C;
  const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < C) {
      float y = (v[i] - mean) * rstd;
      if (mode == 1) {
        y = (1.0f + gv[i]) * y + bv[i];
      } else {
        if (gamma) y = y * gv[i] + bv[i];
      }
      if (act == ACT_LRELU) y = lrelu(y, slope);
      out[(size_t)row * ldo + c] = y;
    }
  }
}

// ALBERT embeddings: LN_128(word[id] + type[0] + pos[t]), eps 1e-12 (transformers AlbertEmbeddings order: word + type, then + pos)
__global__ void __launch_bounds__(256) albert_embed_kernel(const int* __restrict__ ids, const int* __restrict__ pos, const float* __restrict__ word,
                                                           const float* __restrict__ posw, const float* __restrict__ type0,
                                                           const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ out, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int id = ids[row], ps = pos[row];
  float v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane + i * 32;
    v[i] = (word[id * 128 + c] + type0[c]) + posw[ps * 128 + c];
    s += v[i];
  }
  const float mean = warp_sum(s) / 128.f;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float d = v[i] - mean; q += d * d; }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / 128.f + 1e-12f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = lane + i * 32;
    out[(size_t)row * 128 + c] = (v[i] - mean) * rstd * g[c] + b[c];
  }
}

__global__ void embedding_kernel(const int* __restrict__ ids, const float* __restrict__ table, int C, float* __restrict__ out, int ldo, int rows) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const int id = ids[row];
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[(size_t)row * ldo + c] = table[(size_t)id * C + c];
}

__global__ void fill_style_kernel(const float* __restrict__ style, int ld_style, int style_off, const int* __restrict__ row2seg,
                                  float* __restrict__ out, int ldo, int col_off, int C, int rows) {
  const int row = blockIdx.x;
  if (row >= rows) return;
  const int seg = row2seg[row];
  for (int c = threadIdx.x; c < C; c += blockDim.x) out[(size_t)row * ldo + col_off + c] = style[(size_t)seg * ld_style + style_off + c];
}

// ---------------------------------------------------------------------------------------------------------
// ALBERT self-attention: qkv rows [.., 3*H*64] (q | k | v), one warp per (query row, head).
// scores/softmax exactly two-pass (max, exp, sum) like torch.softmax; 1/sqrt(64) applied to the scores.
// ---------------------------------------------------------------------------------------------------------
constexpr int ATT_MAXK = 16;   // keys per lane: T <= 512
constexpr int ATT_QB = 32;     // queries per CTA (8 warps x 4 queries)
constexpr int ATT_KC = 64;     // keys per shared-memory chunk
constexpr int ATT_LD = 68;     // padded row (floats): 16-B aligned, conflict-free for LDS.128 across rows

// One CTA = 32 queries of one (utterance, head); every warp owns four of them and works on all four at once, so one K or V
// element read from shared memory feeds four FMAs.  K and V pass through one 64-key staging buffer (coalesced loads); the
// scores of the CTA live in shared memory [32][len] between the phases.  Two-pass softmax (max, exp, sum) as torch.softmax,
// every sum in a fixed order: lane = key for the scores, lane = output dimension pair for P V.
// (r01: one CTA per 16 queries with the query and its scores in registers ran at one CTA per SM: 107 us per layer.)
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, float* __restrict__ ctx, const Seg* __restrict__ segs, int heads,
                                                        int ld_s) {
  extern __shared__ __align__(16) float att_smem[];
  float (*kv)[ATT_LD] = reinterpret_cast<float (*)[ATT_LD]>(att_smem);                  // [ATT_KC][ATT_LD]
  float (*qs)[64] = reinterpret_cast<float (*)[64]>(att_smem + ATT_KC * ATT_LD);       // [ATT_QB][64]
  float* S = att_smem + ATT_KC * ATT_LD + ATT_QB * 64;                                  // [ATT_QB][ld_s]
  pdl_trigger();
  pdl_wait();               // before any return: every CTA waits for the previous kernel (common.cuh, programmatic dependent launch)
  const Seg seg = segs[blockIdx.z];
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * ATT_QB;
  const int len = seg.in_len;
  if (q0 >= len) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = heads * 64;
  const int ld = 3 * H;
  const int nchunk = (len + ATT_KC - 1) / ATT_KC;
  for (int idx = tid; idx < ATT_QB * 16; idx += 256) {             // the CTA's queries
    const int r = idx >> 4, c4 = idx & 15;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < len) t = *reinterpret_cast<const float4*>(qkv + (size_t)(seg.in_off + q0 + r) * ld + head * 64 + c4 * 4);
    *reinterpret_cast<float4*>(&qs[r][c4 * 4]) = t;
  }
  auto stage = [&](int chunk, int col_off) {       // rows [chunk*64, +64) of K (col_off = H) or V (2H) -> shared memory
#pragma unroll
    for (int it = 0; it < (ATT_KC * 16) / 256; ++it) {
      const int idx = tid + it * 256;
      const int r = idx >> 4, c4 = idx & 15;
      const int key = chunk * ATT_KC + r;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (key < len) t = *reinterpret_cast<const float4*>(qkv + (size_t)(seg.in_off + key) * ld + col_off + head * 64 + c4 * 4);
      *reinterpret_cast<float4*>(&kv[r][c4 * 4]) = t;
    }
  };
  const int qw = warp * 4;                          // this warp's queries qw .. qw+3 (rows of qs / S)
  // ---- scores: S[q][key] = (q . k) / 8 --------------------------------------------------------------------
  for (int chunk = 0; chunk < nchunk; ++chunk) {
    __syncthreads();
    stage(chunk, H);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < ATT_KC / 32; ++j) {
      const int kl = j * 32 + lane;
      const int key = chunk * ATT_KC + kl;
      if (chunk * ATT_KC + j * 32 < len) {          // warp-uniform
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int d = 0; d < 64; d += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&kv[kl][d]);
          const float4 x0 = *reinterpret_cast<const float4*>(&qs[qw][d]);
          const float4 x1 = *reinterpret_cast<const float4*>(&qs[qw + 1][d]);
          const float4 x2 = *reinterpret_cast<const float4*>(&qs[qw + 2][d]);
          const float4 x3 = *reinterpret_cast<const float4*>(&qs[qw + 3][d]);
          a0 = fmaf(x0.x, t.x, a0); a0 = fmaf(x0.y, t.y, a0); a0 = fmaf(x0.z, t.z, a0); a0 = fmaf(x0.w, t.w, a0);
          a1 = fmaf(x1.x, t.x, a1); a1 = fmaf(x1.y, t.y, a1); a1 = fmaf(x1.z, t.z, a1); a1 = fmaf(x1.w, t.w, a1);
          a2 = fmaf(x2.x, t.x, a2); a2 = fmaf(x2.y, t.y, a2); a2 = fmaf(x2.z, t.z, a2); a2 = fmaf(x2.w, t.w, a2);
          a3 = fmaf(x3.x, t.x, a3); a3 = fmaf(x3.y, t.y, a3); a3 = fmaf(x3.z, t.z, a3); a3 = fmaf(x3.w, t.w, a3);
        }
        if (key < len) {
          S[(qw + 0) * ld_s + key] = a0 * 0.125f; S[(qw + 1) * ld_s + key] = a1 * 0.125f;
          S[(qw + 2) * ld_s + key] = a2 * 0.125f; S[(qw + 3) * ld_s + key] = a3 * 0.125f;
        }
      }
    }
  }
  __syncwarp();
  // ---- softmax of the warp's four rows: lane i owns keys i, i+32, ... (ascending), then a butterfly over the lanes -----
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) {
    float* row = S + (qw + qq) * ld_s;
    float mx = -INFINITY;
    for (int k = lane; k < len; k += 32) mx = fmaxf(mx, row[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < len; k += 32) {
      const float e = expf(row[k] - mx);
      row[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int k = lane; k < len; k += 32) row[k] *= inv;
  }
  // ---- P V: lane = output dimensions (lane, lane+32), keys in ascending order ---------------------------------
  float o[4][2];
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) { o[qq][0] = 0.f; o[qq][1] = 0.f; }
  for (int chunk = 0; chunk < nchunk; ++chunk) {
    __syncthreads();
    stage(chunk, 2 * H);
    __syncthreads();
    const int base = chunk * ATT_KC;
    const int lim = min(ATT_KC, len - base);
    int l = 0;
    for (; l + 4 <= lim; l += 4) {
      float4 pq[4];
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) pq[qq] = *reinterpret_cast<const float4*>(S + (qw + qq) * ld_s + base + l);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v0 = kv[l + i][lane], v1 = kv[l + i][lane + 32];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          const float pk = i == 0 ? pq[qq].x : i == 1 ? pq[qq].y : i == 2 ? pq[qq].z : pq[qq].w;
          o[qq][0] = fmaf(pk, v0, o[qq][0]);
          o[qq][1] = fmaf(pk, v1, o[qq][1]);
        }
      }
    }
    for (; l < lim; ++l) {
      const float v0 = kv[l][lane], v1 = kv[l][lane + 32];
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const float pk = S[(qw + qq) * ld_s + base + l];
        o[qq][0] = fmaf(pk, v0, o[qq][0]);
        o[qq][1] = fmaf(pk, v1, o[qq][1]);
      }
    }
  }
#pragma unroll
  for (int qq = 0; qq < 4; ++qq) {
    const int qi = q0 + qw + qq;
    if (qi < len) {
      float* orow = ctx + (size_t)(seg.in_off + qi) * H + head * 64;
      orow[lane] = o[qq][0];
      orow[lane + 32] = o[qq][1];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// duration head: sigmoid -> sum over bins -> / speed -> round-half-even -> clamp(min 1)   (kmodel.py:98-103,115)
// ---------------------------------------------------------------------------------------------------------
__global__ void duration_kernel(const float* __restrict__ logits, int ld, int nbins, const int* __restrict__ row2seg,
                                const float* __restrict__ speeds, float* __restrict__ duration, int* __restrict__ pred_dur, int rows) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float s = 0.f;
  for (int k = 0; k < nbins; ++k) s += sigmoidf_acc(logits[(size_t)row * ld + k]);
  const float d = __fdiv_rn(s, speeds[row2seg[row]]);
  duration[row] = d;
  const float r = rintf(d);            // round half to even == torch.round
  // sum of 50 sigmoids / speed: at most 50 000 with the speeds begin() admits; the upper clamp only guards the scan against a
  // NaN / inf logit (float -> int conversion saturates), it never changes a representable duration
  int n = (r >= 65536.f) ? 65536 : (int)r;
  if (!(r >= 1.f)) n = 1;
  pred_dur[row] = n;
}

// per-utterance inclusive scan of the integer durations (T <= 512 per utterance)
__global__ void __launch_bounds__(512) dur_scan_kernel(const int* __restrict__ pred_dur, const Seg* __restrict__ segs, int* __restrict__ cum,
                                                       int* __restrict__ totals) {
  __shared__ int buf[512];
  const Seg seg = segs[blockIdx.x];
  const int t = threadIdx.x;
  int carry = 0;
  for (int base = 0; base < seg.in_len; base += 512) {
    const int i = base + t;
    int v = (i < seg.in_len) ? pred_dur[seg.in_off + i] : 0;
    buf[t] = v;
    __syncthreads();
    for (int o = 1; o < 512; o <<= 1) {
      int add = (t >= o) ? buf[t - o] : 0;
      __syncthreads();
      buf[t] += add;
      __syncthreads();
    }
    if (i < seg.in_len) cum[seg.in_off + i] = carry + buf[t];
    carry += buf[511];
    __syncthreads();
  }
  if (t == 0) totals[blockIdx.x] = carry;
}

// alignment index map: indices[f] = token whose run covers frame f  (== repeat_interleave(arange(T), pred_dur), kmodel.py:116-119)
// segs here: in = token rows, out = frame rows.
__global__ void indices_kernel(const int* __restrict__ cum, const Seg* __restrict__ segs, int* __restrict__ indices, int total_frames,
                               const int* __restrict__ frame2seg) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= total_frames) return;
  const Seg seg = segs[frame2seg[f]];
  const int lf = f - seg.out_off;
  const int* c = cum + seg.in_off;
  int lo = 0, hi = seg.in_len - 1;      // first token with cum[token] > lf
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (c[mid] > lf) hi = mid; else lo = mid + 1;
  }
  indices[f] = lo;
}

// length regulation: dst[f, :] = src[token_row(indices[f]), :]   (the one-hot matmuls of kmodel.py:132,139 are this gather)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, int lds, const int* __restrict__ indices,
                                                          const int* __restrict__ frame2seg, const Seg* __restrict__ segs,
                                                          float* __restrict__ dst, int ldd, int C4, int rows) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const Seg seg = segs[frame2seg[row]];
  const float4* s = reinterpret_cast<const float4*>(src + (size_t)(seg.in_off + indices[row]) * lds);
  float4* d = reinterpret_cast<float4*>(dst + (size_t)row * ldd);
  for (int c = lane; c < C4; c += 32) d[c] = s[c];
}

}  // namespace

cudaError_t launch_layernorm(const float* x, int ldx, const float* res, int ldr, float* out, int ldo, int rows, int C, float eps,
                             int mode, const float* gamma, const float* beta, const float* gb, int ld_gb, const int* row2seg,
                             int act, float slope, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  if (C > 32 * MAXV) return cudaErrorInvalidValue;
  cudaError_t le = launch_pdl(layernorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, x, ldx, res, ldr, out, ldo, rows, C, eps, mode, gamma, beta, gb, ld_gb, row2seg,
                              act, slope);
  if (le != cudaSuccess) return le;
  return launched();
}
cudaError_t launch_albert_embed(const int* ids, const int* pos, const float* word, const float* posw, const float* type0,
                                const float* g, const float* b, float* out, int rows, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  albert_embed_kernel<<<(rows + 7) / 8, 256, 0, s>>>(ids, pos, word, posw, type0, g, b, out, rows);
  return launched();
}
cudaError_t launch_embedding(const int* ids, const float* table, int C, float* out, int ldo, int rows, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  embedding_kernel<<<rows, 128, 0, s>>>(ids, table, C, out, ldo, rows);
  return launched();
}
cudaError_t launch_fill_style(const float* style, int ld_style, int style_off, const int* row2seg, float* out, int ldo, int col_off, int C,
                              int rows, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  fill_style_kernel<<<rows, 128, 0, s>>>(style, ld_style, style_off, row2seg, out, ldo, col_off, C, rows);
  return launched();
}
cudaError_t launch_attention(const float* qkv, float* ctx, const Seg* segs, int nseg, int max_len, int heads, cudaStream_t s) {
  if (nseg <= 0 || max_len <= 0) return cudaSuccess;
  if (max_len > 32 * ATT_MAXK) return cudaErrorInvalidValue;
  const int ld_s = (max_len + 3) & ~3;                      // float4 reads of the probabilities
  const size_t smem = sizeof(float) * ((size_t)ATT_KC * ATT_LD + (size_t)ATT_QB * 64 + (size_t)ATT_QB * ld_s);
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  dim3 grid((max_len + ATT_QB - 1) / ATT_QB, heads, nseg);
  cudaError_t le = launch_pdl(attention_kernel, grid, dim3(256), smem, s, qkv, ctx, segs, heads, ld_s);
  if (le != cudaSuccess) return le;
  return launched();
}
cudaError_t launch_duration(const float* logits, int ld, int nbins, const int* row2seg, const float* speeds, float* duration, int* pred_dur,
                            int rows, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  duration_kernel<<<(rows + 127) / 128, 128, 0, s>>>(logits, ld, nbins, row2seg, speeds, duration, pred_dur, rows);
  return launched();
}
cudaError_t launch_dur_scan(const int* pred_dur, const Seg* segs, int nseg, int* cum, int* totals, cudaStream_t s) {
  if (nseg <= 0) return cudaSuccess;
  dur_scan_kernel<<<nseg, 512, 0, s>>>(pred_dur, segs, cum, totals);
  return launched();
}
cudaError_t launch_indices(const int* cum, const Seg* segs, int nseg, int* indices, int total_frames, const int* frame2seg, cudaStream_t s) {
  if (total_frames <= 0) return cudaSuccess;
  indices_kernel<<<(total_frames + 255) / 256, 256, 0, s>>>(cum, segs, indices, total_frames, frame2seg);
  return launched();
}
cudaError_t launch_gather_rows(const float* src, int lds, const int* indices, const int* frame2seg, const Seg* segs, float* dst, int ldd,
                               int C, int rows, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  if ((C & 3) || (lds & 3) || (ldd & 3)) return cudaErrorInvalidValue;
  gather_rows_kernel<<<(rows + 7) / 8, 256, 0, s>>>(src, lds, indices, frame2seg, segs, dst, ldd, C / 4, rows);
  return launched();
}

// style rows of resident voice packs: dst[b, :] = *src[b]  (src[b] = device address of pack[slot_b][row_b])
__global__ void gather_ptr_rows_kernel(const float* const* __restrict__ src, float* __restrict__ dst, int C) {
  const float* row = src[blockIdx.x];
  for (int c = threadIdx.x; c < C; c += blockDim.x) dst[(size_t)blockIdx.x * C + c] = row[c];
}
cudaError_t launch_gather_ptr_rows(const float* const* src, float* dst, int C, int rows, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  gather_ptr_rows_kernel<<<rows, 256, 0, s>>>(src, dst, C);
  return launched();
}

}  // namespace ttsb
