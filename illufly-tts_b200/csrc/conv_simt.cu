// fp32 implicit-GEMM 1-D convolution over ragged channels-last rows (CUDA cores, FFMA).
// This is the strict-precision path: every Linear / Conv1d / polyphase ConvTranspose1d of the graph whose result must
// track the fp32 reference (the whole text side, which decides the integer durations) runs here; the vocoder convs run
// here in fp32 mode and on the tcgen05 kernel (conv_umma.cu) in bf16 mode.
// Tile: 128 output rows x 128 output channels per CTA, K-chunks of 16 input channels per tap, 256 threads, 8x8 per thread,
// register-prefetch double buffering.  AdaIN(+LeakyReLU | +Snake) is applied to the operand while it is staged, bias /
// residual / scale / activation / running-sum and the next AdaIN's statistics are applied in the epilogue.
#include "common.cuh"

namespace ttsb {

namespace {
constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int LDA = BM + 4;

struct __align__(16) F4 { float v[4]; };

__global__ void __launch_bounds__(NT, 2) conv_simt_kernel(const ConvArgs p) {
  extern __shared__ float dyn[];            // [3][CinR]: mean, a, b  (+ alpha, 1/alpha when snake) -> 5*CinR
  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const Tile tile = p.tiles[blockIdx.x];
  const Seg seg = p.segs[tile.seg];
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int CinR = (p.Cin + 3) & ~3;
  float* s_mean = dyn;
  float* s_a = dyn + CinR;
  float* s_b = dyn + 2 * CinR;
  float* s_al = dyn + 3 * CinR;
  float* s_ial = dyn + 4 * CinR;

  if (p.pro >= PRO_ADAIN_LRELU) {
    for (int c = tid; c < p.Cin; c += NT) {
      const double* st = p.pstats + ((size_t)tile.seg * p.ld_pstats + c) * 2;
      float g = p.gb[(size_t)tile.seg * p.ld_gb + c];
      float b = p.gb[(size_t)tile.seg * p.ld_gb + p.beta_off + c];
      float mean, a, bb;
      adain_coef(st, seg.in_len, g, b, mean, a, bb);
      s_mean[c] = mean; s_a[c] = a; s_b[c] = bb;
      if (p.pro == PRO_ADAIN_SNAKE) { float al = p.alpha[c]; s_al[c] = al; s_ial[c] = 1.0f / al; }
    }
  }
  __syncthreads();

  const int kchunks = (p.Cin + BK - 1) / BK;
  const int iters = p.k * kchunks;

  // A-stage mapping: 512 float4 per stage -> 2 per thread; idx -> (row = idx/4, c4 = idx%4)
  // B-stage mapping: 16 rows x 32 float4 -> 2 per thread; idx -> (kk = idx/32, n4 = idx%32)
  F4 ra[2], rb[2];

  auto load_stage = [&](int it) {
    const int j = it / kchunks;
    const int ci0 = (it - j * kchunks) * BK;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const int idx = tid + l * NT;
      const int r = idx >> 2, c4 = idx & 3;
      const int t = tile.t0 + r;
      const int ci = ci0 + c4 * 4;
      F4 v; v.v[0] = v.v[1] = v.v[2] = v.v[3] = 0.f;
      const int tin = t * p.stride + j * p.dil - p.pad;
      const int tr = tin >> p.in_shr;
      if (t < seg.out_len && tin >= 0 && tr < seg.in_len && ci < p.Cin) {
        const float* src = p.x + (size_t)(seg.in_off + tr) * p.ldx + ci;
        if ((p.ldx & 3) == 0) {
          v = *reinterpret_cast<const F4*>(src);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (ci + e < p.Cin) v.v[e] = src[e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = ci + e;
          float x = v.v[e];
          if (c >= p.Cin) { x = 0.f; }
          else if (p.pro == PRO_LRELU) { x = lrelu(x, p.slope); }
          else if (p.pro == PRO_ADAIN_LRELU) { x = lrelu(fmaf(x - s_mean[c], s_a[c], s_b[c]), p.slope); }
          else if (p.pro == PRO_ADAIN_SNAKE) { x = snake(fmaf(x - s_mean[c], s_a[c], s_b[c]), s_al[c], s_ial[c]); }
          v.v[e] = x;
        }
      }
      ra[l] = v;
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const int idx = tid + l * NT;
      const int kk = idx >> 5, n4 = idx & 31;
      const int ci = ci0 + kk;
      const int n = n0 + n4 * 4;
      F4 v; v.v[0] = v.v[1] = v.v[2] = v.v[3] = 0.f;
      if (ci < p.Cin && n < p.ldw) v = *reinterpret_cast<const F4*>(p.w + ((size_t)j * p.Cin + ci) * p.ldw + n);
      rb[l] = v;
    }
  };
  auto store_stage = [&](int buf) {
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const int idx = tid + l * NT;
      const int r = idx >> 2, c4 = idx & 3;
#pragma unroll
      for (int e = 0; e < 4; ++e) As[buf][c4 * 4 + e][r] = ra[l].v[e];
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const int idx = tid + l * NT;
      const int kk = idx >> 5, n4 = idx & 31;
      *reinterpret_cast<F4*>(&Bs[buf][kk][n4 * 4]) = rb[l];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int tx = tid & 15, ty = tid >> 4;   // cols tx*4 + {0..3} and 64 + tx*4 + {0..3}; rows ty*4.. and 64 + ty*4..

  load_stage(0);
  store_stage(0);
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) load_stage(it + 1);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const F4 a0 = *reinterpret_cast<const F4*>(&As[buf][kk][ty * 4]);
      const F4 a1 = *reinterpret_cast<const F4*>(&As[buf][kk][64 + ty * 4]);
      const F4 b0 = *reinterpret_cast<const F4*>(&Bs[buf][kk][tx * 4]);
      const F4 b1 = *reinterpret_cast<const F4*>(&Bs[buf][kk][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a0.v[i], b0.v[j], acc[i][j]);
          acc[i][j + 4] = fmaf(a0.v[i], b1.v[j], acc[i][j + 4]);
          acc[i + 4][j] = fmaf(a1.v[i], b0.v[j], acc[i + 4][j]);
          acc[i + 4][j + 4] = fmaf(a1.v[i], b1.v[j], acc[i + 4][j + 4]);
        }
      }
    }
    if (it + 1 < iters) {
      store_stage(buf ^ 1);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  float ps[8], pq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ps[j] = 0.f; pq[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = (i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4));
    const int t = tile.t0 + r;
    if (t >= seg.out_len) continue;
    const size_t orow = (size_t)(seg.out_off + t);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int cl = h * 64 + tx * 4;
      const int co = n0 + cl;
      if (co >= p.Cout) continue;
      float v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float x = acc[i][h * 4 + e];
        const int c = co + e;
        if (c < p.Cout) {
          if (p.bias) x += p.bias[c];
          if (p.res) x += p.res[orow * p.ldr + c];
          if (p.act == ACT_GELU_NEW) x = gelu_new(x);
          else if (p.act == ACT_LRELU) x = lrelu(x, p.act_slope);
          x *= p.out_scale;
          if (p.accumulate) x += p.y[orow * p.ldy + c];
        } else x = 0.f;
        v[e] = x;
      }
      float* dst = p.y + orow * p.ldy + co;
      if (((p.ldy & 3) == 0) && co + 3 < p.Cout && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        F4 o; o.v[0] = v[0]; o.v[1] = v[1]; o.v[2] = v[2]; o.v[3] = v[3];
        *reinterpret_cast<F4*>(dst) = o;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (co + e < p.Cout) dst[e] = v[e];
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { ps[h * 4 + e] += v[e]; pq[h * 4 + e] += v[e] * v[e]; }
    }
  }
  if (p.ostats) {
    // deterministic in-CTA reduction: every thread parks its partials (the operand tiles are dead by now), then one
    // thread per channel adds the 16 row-groups in a fixed order; the cross-CTA sum is a double atomic (order changes
    // it by ~1e-16 relative, below fp32 resolution of the statistics)
    float* part_s = &As[0][0][0];     // [16][BN]
    float* part_q = &Bs[0][0][0];     // [16][BN]
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int cl = h * 64 + tx * 4 + e;
        part_s[ty * BN + cl] = ps[h * 4 + e];
        part_q[ty * BN + cl] = pq[h * 4 + e];
      }
    __syncthreads();
    if (tid < BN && n0 + tid < p.Cout) {
      float s = 0.f, q = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) { s += part_s[g * BN + tid]; q += part_q[g * BN + tid]; }
      double* st = p.ostats + ((size_t)tile.seg * p.ld_ostats + n0 + tid) * 2;
      atomicAdd(st, (double)s);
      atomicAdd(st + 1, (double)q);
    }
  }
}

// Style projection: out[b][n] = bias[n] + sum_k style[b][k] W[k][n] for a HANDFUL of rows b (one per utterance) and tens of thousands
// of columns n (every AdaIN / AdaLayerNorm `fc` of a style half side by side).  One thread per column, the style rows transposed in
// shared memory ([k][32 utterances]: one 16-B load serves four utterances), weight reads coalesced over n and issued eight k ahead.
// Same arithmetic as conv_simt_kernel gives for this shape - acc = 0, acc = fmaf(x_k, w_k, acc) for ascending k, + bias - so the
// outputs are bit-identical (they become the scales and shifts that decide the integer durations: tested).  The 128 x 128-tile
// kernel spent 82 us on it with a quarter of its rows valid at batch 32 (and one row at batch 1), twice per step.
constexpr int SP_K = 128;
__global__ void __launch_bounds__(256) style_proj_kernel(const float* __restrict__ style, int ld_style, int nseg, const float* __restrict__ w, int ldw,
                                                         const float* __restrict__ bias, float* __restrict__ out, int ldo, int K, int N) {
  __shared__ __align__(16) float s_t[SP_K][32];
  const int n = blockIdx.x * 256 + threadIdx.x;
  const int nc = n < N ? n : N - 1;                       // clamped: every load is valid, the store is guarded
  for (int b0 = 0; b0 < nseg; b0 += 32) {
    const int nb = min(32, nseg - b0);
    for (int i = threadIdx.x; i < SP_K * 32; i += 256) {
      const int k = i >> 5, b = i & 31;
      s_t[k][b] = (k < K && b < nb) ? style[(size_t)(b0 + b) * ld_style + k] : 0.f;
    }
    __syncthreads();
    float acc[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) acc[b] = 0.f;
    for (int k0 = 0; k0 < K; k0 += 8) {
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = (k0 + j < K) ? w[(size_t)(k0 + j) * ldw + nc] : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (k0 + j < K) {
#pragma unroll
          for (int b4 = 0; b4 < 8; ++b4) {
            const float4 s4 = *reinterpret_cast<const float4*>(&s_t[k0 + j][4 * b4]);
            acc[4 * b4 + 0] = fmaf(s4.x, wv[j], acc[4 * b4 + 0]);
            acc[4 * b4 + 1] = fmaf(s4.y, wv[j], acc[4 * b4 + 1]);
            acc[4 * b4 + 2] = fmaf(s4.z, wv[j], acc[4 * b4 + 2]);
            acc[4 * b4 + 3] = fmaf(s4.w, wv[j], acc[4 * b4 + 3]);
          }
        }
      }
    }
    if (n < N) {
      const float bn = bias ? bias[n] : 0.f;
#pragma unroll
      for (int b = 0; b < 32; ++b)
        if (b < nb) out[(size_t)(b0 + b) * ldo + n] = bias ? acc[b] + bn : acc[b];
    }
    __syncthreads();
  }
}

// y[r] = bias + sum_c x[r][c] w[c]: a Conv1d(C, 1, 1) (F0_proj, N_proj) over every row of a space.  A warp stages 32 rows x 32
// channels through padded shared memory (coalesced 128-B row segments in, conflict-free columns out) and each lane then walks ITS
// row with one accumulator, channels ascending - the order of conv_simt_kernel, so F0 and N are bit-identical to what the tiled
// kernel gave (F0 feeds the phase integrator: tested).  The tiled kernel computed 128 output columns to keep one: 82 us per launch.
__global__ void __launch_bounds__(128) rowdot_kernel(const float* __restrict__ x, int ldx, int rows, int C, const float* __restrict__ w, int ldw,
                                                     const float* __restrict__ bias, float* __restrict__ y, int ldy) {
  __shared__ float tile[4][32][33];
  __shared__ float ws[512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < C; c += 128) ws[c] = w[(size_t)c * ldw];
  __syncthreads();
  const int r0 = (blockIdx.x * 4 + warp) * 32;
  if (r0 >= rows) return;
  float acc = 0.f;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int cc = min(c0 + lane, C - 1);
#pragma unroll
    for (int i = 0; i < 32; ++i) tile[warp][i][lane] = x[(size_t)min(r0 + i, rows - 1) * ldx + cc];
    __syncwarp();
    const int nj = min(32, C - c0);
    if (nj == 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc = fmaf(tile[warp][lane][j], ws[c0 + j], acc);
    } else {
      for (int j = 0; j < nj; ++j) acc = fmaf(tile[warp][lane][j], ws[c0 + j], acc);
    }
    __syncwarp();
  }
  if (r0 + lane < rows) y[(size_t)(r0 + lane) * ldy] = bias ? acc + bias[0] : acc;
}
}  // namespace

cudaError_t launch_style_proj(const float* style, int ld_style, int nseg, const float* w, int ldw, const float* bias, float* out, int ldo, int K,
                              int N, cudaStream_t s) {
  if (nseg <= 0 || N <= 0) return cudaSuccess;
  if (K <= 0 || K > SP_K) return cudaErrorInvalidValue;
  style_proj_kernel<<<(N + 255) / 256, 256, 0, s>>>(style, ld_style, nseg, w, ldw, bias, out, ldo, K, N);
  return launched();
}

cudaError_t launch_rowdot(const float* x, int ldx, int rows, int C, const float* w, int ldw, const float* bias, float* y, int ldy, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  if (C <= 0 || C > 512) return cudaErrorInvalidValue;
  rowdot_kernel<<<(rows + 127) / 128, 128, 0, s>>>(x, ldx, rows, C, w, ldw, bias, y, ldy);
  return launched();
}

cudaError_t launch_conv_simt(const ConvArgs& a, cudaStream_t s) {
  if (a.ntiles <= 0) return cudaSuccess;
  const int CinR = (a.Cin + 3) & ~3;
  const size_t dyn = (a.pro >= PRO_ADAIN_LRELU) ? (size_t)5 * CinR * sizeof(float) : 0;
  static bool attr_set[64] = {false};          // cudaFuncSetAttribute is per device
  int dev = 0;
  cudaError_t ed = cudaGetDevice(&dev);
  if (ed != cudaSuccess) return ed;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  dim3 grid(a.ntiles, (a.Cout + BN - 1) / BN);
  conv_simt_kernel<<<grid, NT, dyn, s>>>(a);
  return launched();
}

}  // namespace ttsb
