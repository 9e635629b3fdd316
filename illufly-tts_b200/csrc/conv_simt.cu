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
}  // namespace

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
