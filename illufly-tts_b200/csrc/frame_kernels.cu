// Frame-side bandwidth-bound kernels: per-utterance channel statistics, AdaIN+Snake, depthwise x2 transposed
// conv, the NSF harmonic source + STFT(20,5), upsample-add, and the fused exp/sin + iSTFT(20,5) overlap-add.
// All of them are single coalesced passes over channels-last rows; the roofline that bounds them is HBM.
#include <stdlib.h>

#include "common.cuh"
#include <curand_kernel.h>
#include <math.h>

namespace ttsb {

__constant__ float c_cos20[20];
__constant__ float c_sin20[20];
__constant__ float c_hann20[20];
static bool g_tables_ready[64] = {false};     // __constant__ memory is per device

static cudaError_t ensure_tables() {
  int dev = 0;
  cudaError_t ed = cudaGetDevice(&dev);
  if (ed != cudaSuccess) return ed;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (g_tables_ready[dev]) return cudaSuccess;
  float c[20], s[20], w[20];
  for (int i = 0; i < 20; ++i) {
    const double a = 2.0 * M_PI * (double)i / 20.0;
    c[i] = (float)cos(a);
    s[i] = (float)sin(a);
    w[i] = (float)(0.5 * (1.0 - cos(a)));   // periodic Hann(20)
  }
  c[5] = 0.f; c[15] = 0.f; s[0] = 0.f; s[10] = 0.f;   // exact zeros of the 20-point twiddles
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(c_cos20, c, sizeof(c))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_sin20, s, sizeof(s))) != cudaSuccess) return e;
  if ((e = cudaMemcpyToSymbol(c_hann20, w, sizeof(w))) != cudaSuccess) return e;
  g_tables_ready[dev] = true;
  return cudaSuccess;
}

namespace {

// ---------------------------------------------------------------------------------------------------------
// per-(utterance, channel) sum and sum of squares over time (double), for InstanceNorm
// grid (ceil(C/128), row splits, nseg), block (32, 8): a thread owns 4 channels (one 16-B load per row when aligned)
// ---------------------------------------------------------------------------------------------------------
constexpr int CS_ROWS = 512;    // rows per block
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ x, int ldx, int C, const Seg* __restrict__ segs,
                                                       double* __restrict__ stats, int ld_stats) {
  __shared__ double sh[2][8][128];
  const Seg seg = segs[blockIdx.z];
  const int c0 = blockIdx.x * 128 + threadIdx.x * 4;
  const int ty = threadIdx.y;
  const int r0 = blockIdx.y * CS_ROWS, r1 = min(seg.in_len, r0 + CS_ROWS);
  if (r0 >= seg.in_len) return;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vec = ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && c0 + 3 < C;
  for (int r = r0 + ty; r < r1; r += 8) {
    const float* row = x + (size_t)(seg.in_off + r) * ldx + c0;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec) { const float4 t = *reinterpret_cast<const float4*>(row); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (c0 + e < C) v[e] = row[e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) { s[e] += v[e]; q[e] = fmaf(v[e], v[e], q[e]); }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { sh[0][ty][threadIdx.x * 4 + e] = (double)s[e]; sh[1][ty][threadIdx.x * 4 + e] = (double)q[e]; }
  __syncthreads();
  const int t = ty * 32 + threadIdx.x;          // 256 threads: first 128 reduce the sums, the other 128 the squares
  const int which = t >> 7, cl = t & 127;
  const int c = blockIdx.x * 128 + cl;
  if (c < C) {
    double a = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a += sh[which][i][cl];
    atomicAdd(stats + ((size_t)blockIdx.z * ld_stats + c) * 2 + which, a);
  }
}

__global__ void zero_f64_kernel(double* p, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = 0.0;
}

// ---------------------------------------------------------------------------------------------------------
constexpr int AS_ROWS = 256;    // rows per block
// standalone AdaIN1d + Snake pass (A.8 + A.12): y = v + (1/alpha) sin^2(alpha v), v = (1+g)(x-mean)rstd + b
// 8 B/element algorithmic (statistics come from the producer's epilogue).  grid (row tiles of 32, nseg)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adain_snake_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int C,
                                                          const double* __restrict__ pstats, int ld_pstats, const float* __restrict__ gb, int ld_gb,
                                                          int beta_off, const float* __restrict__ alpha, const Seg* __restrict__ segs) {
  extern __shared__ float coef[];   // [5][C]: A, B + k, 2 alpha A, 2 alpha B, k   (same folding as the conv kernel's producer)
  const Seg seg = segs[blockIdx.y];
  const int r0 = blockIdx.x * AS_ROWS;
  if (r0 >= seg.in_len) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, a, b;
    adain_coef(pstats + ((size_t)blockIdx.y * ld_pstats + c) * 2, seg.in_len, gb[(size_t)blockIdx.y * ld_gb + c],
               gb[(size_t)blockIdx.y * ld_gb + beta_off + c], mean, a, b);
    const float al = alpha[c];
    const float Bf = fmaf(-mean, a, b), k = 0.5f / al;
    coef[c] = a; coef[C + c] = Bf + k; coef[2 * C + c] = 2.f * al * a; coef[3 * C + c] = 2.f * al * Bf; coef[4 * C + c] = k;
  }
  __syncthreads();
  const int C4 = C >> 2;
  const int r1 = min(seg.in_len, r0 + AS_ROWS);
  for (int idx = threadIdx.x; idx < (r1 - r0) * C4; idx += blockDim.x) {
    const int r = r0 + idx / C4, c = (idx % C4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + (size_t)(seg.in_off + r) * ldx + c);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      o[e] = fmaf(-coef[4 * C + c + e], __cosf(fmaf(vv[e], coef[2 * C + c + e], coef[3 * C + c + e])), fmaf(vv[e], coef[c + e], coef[C + c + e]));
    *reinterpret_cast<float4*>(y + (size_t)(seg.out_off + r) * ldy + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// AdainResBlk1d "pool": AdaIN + LeakyReLU, then depthwise ConvTranspose1d(k3,s2,p1,op1): L -> 2L   (A.7)
//   y[2m] = v[m] w1 + b ; y[2m+1] = v[m] w2 + v[m+1] w0 + b       grid (ceil(C/128), ceil(max_out/64), nseg)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pool_up2_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int C,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       const double* __restrict__ pstats, int ld_pstats, const float* __restrict__ gb, int ld_gb,
                                                       int beta_off, float slope, const Seg* __restrict__ segs) {
  const Seg seg = segs[blockIdx.z];
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int t0 = blockIdx.y * 64;
  if (c >= C || t0 >= seg.out_len) return;
  float mean, a, b;
  adain_coef(pstats + ((size_t)blockIdx.z * ld_pstats + c) * 2, seg.in_len, gb[(size_t)blockIdx.z * ld_gb + c],
             gb[(size_t)blockIdx.z * ld_gb + beta_off + c], mean, a, b);
  const float w0 = w[c * 3 + 0], w1 = w[c * 3 + 1], w2 = w[c * 3 + 2], bb = bias[c];
  const int t1 = min(seg.out_len, t0 + 64);
  for (int t = t0; t < t1; ++t) {
    const int m = t >> 1;
    const float vm = lrelu(fmaf(x[(size_t)(seg.in_off + m) * ldx + c] - mean, a, b), slope);
    float o;
    if ((t & 1) == 0) {
      o = vm * w1 + bb;
    } else {
      float vn = 0.f;
      if (m + 1 < seg.in_len) vn = lrelu(fmaf(x[(size_t)(seg.in_off + m + 1) * ldx + c] - mean, a, b), slope);
      o = vm * w2 + vn * w0 + bb;
    }
    y[(size_t)(seg.out_off + t) * ldy + c] = o;
  }
}

// F0_conv / N_conv: weight-normed Conv1d(1,1,k3,s2,p1) on a [2F] curve -> column of a concat buffer (A.10)
__global__ void conv1_s2_kernel(const float* __restrict__ x, float* __restrict__ y, int ldy, const float* __restrict__ w3, const float* __restrict__ bias1,
                                const Seg* __restrict__ segs) {
  const Seg seg = segs[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= seg.out_len) return;
  const float* xs = x + seg.in_off;
  float acc = bias1[0];
  const int i0 = 2 * t - 1;
  if (i0 >= 0) acc = fmaf(w3[0], xs[i0], acc);
  acc = fmaf(w3[1], xs[i0 + 1], acc);
  if (i0 + 2 < seg.in_len) acc = fmaf(w3[2], xs[i0 + 2], acc);
  y[(size_t)(seg.out_off + t) * ldy] = acc;
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int C, int rows) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * C) return;
  const size_t r = i / C; const int c = (int)(i % C);
  dst[r * ldd + c] = src[r * lds + c];
}

// ---------------------------------------------------------------------------------------------------------
// NSF harmonic source (A.11): phase scan.  One thread per (utterance, harmonic): torch.cumsum on CPU accumulates
// fp32 inputs in fp64 and rounds each prefix to fp32, then *2, *pi, *300 in fp32.
// ---------------------------------------------------------------------------------------------------------
// serial form (one thread per utterance and harmonic): the reference order, kept behind TTSB_PHASE_SERIAL=1 for bit-for-bit checks
__global__ void phase_scan_serial_kernel(const float* __restrict__ f0, float* __restrict__ phase, const Seg* __restrict__ segs, int nseg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nseg * 9) return;
  const int sgi = idx / 9, h = idx % 9;
  const Seg seg = segs[sgi];
  const float hm = (float)(h + 1);
  double acc = 0.0;
  for (int d = 0; d < seg.in_len; ++d) {
    const float fn = __fmul_rn(f0[seg.in_off + d], hm);
    const float q = __fdiv_rn(fn, 24000.0f);
    float m = fmodf(q, 1.0f);
    if (m != 0.f && m < 0.f) m = __fadd_rn(m, 1.0f);        // python-style remainder, as c10 does it
    acc += (double)m;
    const float cs = (float)acc;
    float p = __fmul_rn(__fmul_rn(cs, 2.0f), 3.14159274101257324f);
    p = __fmul_rn(p, 300.0f);
    phase[(size_t)(seg.in_off + d) * 9 + h] = p;
  }
}

// One WARP per (utterance, harmonic): the frames in chunks of 32, a shuffle prefix sum of the per-frame terms in double, the
// chunk's carry added on top.  The per-frame term m = fmod(f0 h / 24000, 1) is the same float as in the serial form; the only
// thing that changes is the ORDER of the double additions, and those are exact: m is a float in [0, 1) - a multiple of 2^-32 or
// coarser for any f0 h above ~0.01 Hz - and the partial sums stay below 2^9, so every partial sum is representable in 53 bits and
// association does not matter (below that, a difference of 2^-53 relative could appear, and disappears in the cast to float).
// Bit-for-bit equality with the serial form is a test (test_phase_scan_parallel_equals_serial).  The serial form took 159 us for
// 275 frames - one global-load round trip per frame on 5 of the 148 SMs - at batch 32 and just as long for a single utterance.
__global__ void __launch_bounds__(128) phase_scan_kernel(const float* __restrict__ f0, float* __restrict__ phase, const Seg* __restrict__ segs, int nseg) {
  const int wid = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= nseg * 9) return;
  const int sgi = wid / 9, h = wid % 9;
  const Seg seg = segs[sgi];
  const float hm = (float)(h + 1);
  double carry = 0.0;
  for (int d0 = 0; d0 < seg.in_len; d0 += 32) {
    const int d = d0 + lane;
    float m = 0.f;
    if (d < seg.in_len) {
      const float fn = __fmul_rn(f0[seg.in_off + d], hm);
      const float q = __fdiv_rn(fn, 24000.0f);
      m = fmodf(q, 1.0f);
      if (m != 0.f && m < 0.f) m = __fadd_rn(m, 1.0f);        // python-style remainder, as c10 does it
    }
    double v = (double)m;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += up;
    }
    const double acc = carry + v;
    carry += __shfl_sync(0xffffffffu, v, 31);
    if (d < seg.in_len) {
      const float cs = (float)acc;
      float p = __fmul_rn(__fmul_rn(cs, 2.0f), 3.14159274101257324f);
      p = __fmul_rn(p, 300.0f);
      phase[(size_t)(seg.in_off + d) * 9 + h] = p;
    }
  }
}

// har_source sample n of an utterance (local index), given per-frame phases.  Mirrors F.interpolate(linear,
// align_corners=False) as ATen's CPU kernel evaluates it (probed bit-exact): src = fma(1/300, n+0.5, -0.5),
// out = fma(p[i0], 1-lam, p[i1]*lam).
__device__ __forceinline__ float har_sample(int n, const float* __restrict__ f0, const float* __restrict__ phase, int L,
                                            const float* __restrict__ noise9, const float* __restrict__ lw, float lb,
                                            unsigned long long seed, unsigned long long gidx) {
  const float scale = 0.0033333334140479565f;   // (float)(1.0/300.0)
  float src = __fmaf_rn(scale, (float)n + 0.5f, -0.5f);
  if (src < 0.f) src = 0.f;
  int i0 = (int)floorf(src);
  if (i0 > L - 1) i0 = L - 1;
  float lam1 = src - (float)i0;
  lam1 = fminf(fmaxf(lam1, 0.f), 1.f);
  const float lam0 = 1.0f - lam1;
  const int i1 = i0 + ((i0 < L - 1) ? 1 : 0);
  const float f = f0[n / 300];
  const float uv = (f > 10.0f) ? 1.0f : 0.0f;
  const float namp = (f > 10.0f) ? 0.003f : __fdiv_rn(0.1f, 3.0f);
  float nz[9];
  if (noise9) {
#pragma unroll
    for (int h = 0; h < 9; ++h) nz[h] = noise9[h];
  } else {
    curandStatePhilox4_32_10_t st;
    curand_init(seed, gidx, 0, &st);
    const float4 a = curand_normal4(&st), b = curand_normal4(&st);
    nz[0] = a.x; nz[1] = a.y; nz[2] = a.z; nz[3] = a.w; nz[4] = b.x; nz[5] = b.y; nz[6] = b.z; nz[7] = b.w;
    nz[8] = curand_normal(&st);
  }
  float acc = lb;
#pragma unroll
  for (int h = 0; h < 9; ++h) {
    const float ph = __fmaf_rn(phase[(size_t)i0 * 9 + h], lam0, __fmul_rn(phase[(size_t)i1 * 9 + h], lam1));
    const float sine = __fmul_rn(sinf(ph), 0.1f);
    const float sw = __fadd_rn(__fmul_rn(sine, uv), __fmul_rn(namp, nz[h]));
    acc = fmaf(lw[h], sw, acc);
  }
  return tanhf(acc);
}

// Source + STFT fused: one CTA = 64 STFT frames of one utterance; the 20+5*63 samples (+ reflect halo) they need are
// synthesised into shared memory (and written once to har_src for inspection), then 11 DFT bins per frame.
// har row = [ |X_0..10| , angle(X_0..10) , 0 , 0 ]   (ld 24)
constexpr int SF = 64;
constexpr int SS = 5 * (SF - 1) + 20;   // 335 samples per tile
__global__ void __launch_bounds__(128) source_stft_kernel(const float* __restrict__ f0, const float* __restrict__ noise, const float* __restrict__ lw,
                                                          const float* __restrict__ lb, float* __restrict__ har_src, float* __restrict__ har, int ld_har,
                                                          const float* __restrict__ phase, const Seg* __restrict__ segs, unsigned long long seed,
                                                          const unsigned long long* __restrict__ seg_seeds) {
  __shared__ float xs[SS];
  const Seg seg = segs[blockIdx.y];      // in: L rows of f0 ; out: 60L+1 har rows
  const int fr0 = blockIdx.x * SF;
  if (fr0 >= seg.out_len) return;
  const int L = seg.in_len, N = 300 * L;
  const size_t samp_off = (size_t)seg.in_off * 300;   // utterance's first audio-rate sample
  // in-kernel SineGen noise: with a seed table the stream is keyed by (utterance seed, sample index WITHIN the utterance), so an
  // utterance draws the same noise wherever it sits in a batch; without one (vocoder-only entry) by (seed, position in the buffer)
  const unsigned long long nz_seed = seg_seeds ? seg_seeds[blockIdx.y] : seed;
  const size_t nz_base = seg_seeds ? 0 : samp_off;
  for (int i = threadIdx.x; i < SS; i += blockDim.x) {
    int j = 5 * fr0 + i - 10;            // sample index before reflection
    float v = 0.f;
    const bool inside = (j >= 0 && j < N);
    if (j < 0) j = -j;
    if (j >= N) j = 2 * (N - 1) - j;
    if (j >= 0 && j < N) {
      v = har_sample(j, f0 + seg.in_off, phase + (size_t)seg.in_off * 9, L, noise ? noise + (samp_off + j) * 9 : nullptr, lw, lb[0], nz_seed, nz_base + j);
      if (inside && har_src && i >= 10 && (i < 10 + 5 * SF || 5 * fr0 + i - 10 >= N - 10)) har_src[samp_off + j] = v;
    }
    xs[i] = v;
  }
  __syncthreads();
  for (int f = threadIdx.x; f < SF; f += blockDim.x) {
    const int fr = fr0 + f;
    if (fr >= seg.out_len) break;
    float xw[20];
#pragma unroll
    for (int m = 0; m < 20; ++m) xw[m] = xs[5 * f + m] * c_hann20[m];
    float* out = har + (size_t)(seg.out_off + fr) * ld_har;
#pragma unroll
    for (int k = 0; k <= 10; ++k) {
      float re = 0.f, im = 0.f;
#pragma unroll
      for (int m = 0; m < 20; ++m) {
        const int tw = (k * m) % 20;
        re = fmaf(xw[m], c_cos20[tw], re);
        im = fmaf(-xw[m], c_sin20[tw], im);
      }
      if (k == 0 || k == 10) im = 0.f;   // r2c transforms return +0 there
      out[k] = hypotf(re, im);
      out[11 + k] = atan2f(im, re);
    }
    out[22] = 0.f; out[23] = 0.f;
  }
}

// x = ups(x) [flat polyphase buffer, shifted by the transposed conv's padding] (+ ReflectionPad1d((1,0))) + x_source ; stats
// grid (ceil(C/128), ceil(max_out/64), nseg)
__global__ void __launch_bounds__(128) up_add_kernel(const float* __restrict__ up, int ldu, int shift, int reflect, const float* __restrict__ xs, int ldxs,
                                                     float* __restrict__ y, int ldy, int C, const Seg* __restrict__ segs, double* __restrict__ ostats,
                                                     int ld_ostats) {
  const Seg seg = segs[blockIdx.z];
  const int c = blockIdx.x * 128 + threadIdx.x;
  const int t0 = blockIdx.y * 64;
  if (c >= C || t0 >= seg.out_len) return;
  const int t1 = min(seg.out_len, t0 + 64);
  float s = 0.f, q = 0.f;
  for (int t = t0; t < t1; ++t) {
    const int tu = reflect ? (t == 0 ? 1 : t - 1) : t;
    const float v = up[(size_t)(seg.in_off + shift + tu) * ldu + c] + xs[(size_t)(seg.out_off + t) * ldxs + c];
    y[(size_t)(seg.out_off + t) * ldy + c] = v;
    s += v; q += v * v;
  }
  if (ostats) {
    double* st = ostats + ((size_t)blockIdx.z * ld_ostats + c) * 2;
    atomicAdd(st, (double)s);
    atomicAdd(st + 1, (double)q);
  }
}

// ---------------------------------------------------------------------------------------------------------
// conv_post epilogue + iSTFT (A.11): spec = exp(x[:11]), phase = sin(x[11:]), z = spec e^{j phase};
// y = overlap-add(irfft20(z) * hann) / sum(hann^2), centre-trimmed.
// One thread = one STFT frame: its 22 inputs come in with six 16-B loads, the 20 windowed irfft outputs are produced in
// registers with compile-time twiddles (the even / odd symmetry of cos / sin halves the work) and parked in shared memory;
// then the block writes 5 output samples per frame, each the sum of the 4 frames that overlap it, coalesced.
// Algorithmic bytes: read 22 fp32 per frame, write 5 fp32 per frame  (12.96 kB per decoder frame F): HBM bound.
// ---------------------------------------------------------------------------------------------------------
constexpr int IS_FPB = 256;                 // frames (= 1280 output samples) per block
constexpr int IS_HALO = 3;                  // a sample also needs the 3 frames before its own
constexpr int IS_LD = 21;                   // padded frame row in shared memory

struct Tw20 {                               // cos / sin(2 pi i / 20) as compile-time constants
  __host__ __device__ static constexpr float c(int i) {
    constexpr float t[20] = {1.0f, 0.951056516295153572f, 0.809016994374947424f, 0.587785252292473129f, 0.309016994374947424f, 0.0f,
                             -0.309016994374947424f, -0.587785252292473129f, -0.809016994374947424f, -0.951056516295153572f, -1.0f,
                             -0.951056516295153572f, -0.809016994374947424f, -0.587785252292473129f, -0.309016994374947424f, 0.0f,
                             0.309016994374947424f, 0.587785252292473129f, 0.809016994374947424f, 0.951056516295153572f};
    return t[i % 20];
  }
  __host__ __device__ static constexpr float s(int i) { return c((i + 15) % 20); }   // sin(x) = cos(x - pi/2)
  __host__ __device__ static constexpr float w(int m) { return 0.5f * (1.0f - c(m)); }   // periodic Hann(20)
};

// PRECISE: expf / sinf / sincosf instead of the fast intrinsics, whose absolute error grows with |x| (conv_post's log-magnitude and
// phase logits are unbounded): the strict fp32 tier uses it, the 16-bit operand modes keep the intrinsics
template <bool PRECISE>
__global__ void __launch_bounds__(IS_FPB + 32) istft_kernel(const float* __restrict__ cp, int ld, float* __restrict__ audio, const Seg* __restrict__ segs) {
  __shared__ float fr_out[(IS_FPB + IS_HALO) * IS_LD];
  const Seg seg = segs[blockIdx.y];        // in: frames (120F+1), out: samples (600F)
  const int f0 = blockIdx.x * IS_FPB;      // this block owns padded-signal samples [5 f0, 5 f0 + 1280)
  if (5 * f0 - 10 >= seg.out_len) return;
  const int tid = threadIdx.x;
  if (tid < IS_FPB + IS_HALO) {
    const int fr = f0 - IS_HALO + tid;
    float* o = fr_out + tid * IS_LD;
    if (fr >= 0 && fr < seg.in_len) {
      const float4* row = reinterpret_cast<const float4*>(cp + (size_t)(seg.in_off + fr) * ld);
      float x[24];
#pragma unroll
      for (int i = 0; i < 6; ++i) { const float4 t = row[i]; x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w; }
      float re[11], im[11];
#pragma unroll
      for (int k = 0; k <= 10; ++k) {
        const float mag = PRECISE ? expf(x[k]) : __expf(x[k]);
        float sn, cs;
        if (PRECISE) sincosf(sinf(x[11 + k]), &sn, &cs);
        else __sincosf(__sinf(x[11 + k]), &sn, &cs);
        re[k] = mag * cs; im[k] = mag * sn;
      }
      // x[m] = (X0 + (-1)^m X10 + 2 sum_{k=1..9} (Re_k cos t - Im_k sin t)) / 20,  t = 2 pi k m / 20.
      // Two symmetries: m -> 20-m flips the sine part (x[m] = C - S, x[20-m] = C + S), and m -> 10-m multiplies the bin k terms
      // by (-1)^k (cos) and -(-1)^k (sin): with the even-k and odd-k partial sums Ce, Co, Se, So of m = 0..5,
      //   C[m] = Ce + Co, C[10-m] = Ce - Co, S[m] = Se + So, S[10-m] = So - Se.
      // 18 FMAs for four outputs instead of 18 for two (the kernel is instruction-issue bound: ncu, 83 % issue slots busy);
      // twiddles that are exactly zero are dropped at compile time.
#pragma unroll
      for (int m = 0; m <= 5; ++m) {
        float Ce = re[0] + ((m & 1) ? -re[10] : re[10]), Co = 0.f, Se = 0.f, So = 0.f;
#pragma unroll
        for (int k = 1; k <= 9; ++k) {
          const float c2 = 2.0f * Tw20::c(k * m), s2 = 2.0f * Tw20::s(k * m);
          if (k & 1) {
            if (c2 != 0.0f) Co = fmaf(c2, re[k], Co);
            if (s2 != 0.0f) So = fmaf(s2, im[k], So);
          } else {
            if (c2 != 0.0f) Ce = fmaf(c2, re[k], Ce);
            if (s2 != 0.0f) Se = fmaf(s2, im[k], Se);
          }
        }
        const float C = Ce + Co, S = Se + So;
        o[m] = (C - S) * (0.05f * Tw20::w(m));
        if (m >= 1) o[20 - m] = (C + S) * (0.05f * Tw20::w(20 - m));
        if (m <= 4) {
          const float Cm = Ce - Co, Sm = So - Se;
          o[10 - m] = (Cm - Sm) * (0.05f * Tw20::w(10 - m));
          if (m >= 1) o[10 + m] = (Cm + Sm) * (0.05f * Tw20::w(10 + m));
        }
      }
    } else {
#pragma unroll
      for (int m = 0; m < 20; ++m) o[m] = 0.f;
    }
  }
  __syncthreads();
  // 1280 samples of the padded signal, index p; output sample n = p - 10
  for (int i = tid; i < 5 * IS_FPB; i += blockDim.x) {
    const int p = 5 * f0 + i;
    const int n = p - 10;
    if (n < 0 || n >= seg.out_len) continue;
    const int fl = i / 5 + IS_HALO;          // local index of frame floor(p / 5)
    const int r = i - 5 * (i / 5);
    const int frn = f0 + i / 5;
    if (frn >= 3 && frn < seg.in_len) {
      // all four frames that overlap this sample exist: the squared periodic Hann(20) windows at hop 5 sum to exactly 3/2
      // (every r), so the envelope division is one multiplication - no table loads, no divide, in 99.99 % of the samples
      const float* q = fr_out + fl * IS_LD + r;
      const float acc4 = (q[0] + q[5 - IS_LD]) + (q[10 - 2 * IS_LD] + q[15 - 3 * IS_LD]);
      audio[(size_t)seg.out_off + n] = acc4 * 0.666666686534881591796875f;
      continue;
    }
    float acc = 0.f, env = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int fr = f0 + i / 5 - d;
      if (fr >= 0 && fr < seg.in_len) {
        const int m = r + 5 * d;
        acc += fr_out[(fl - d) * IS_LD + m];
        const float w = 0.5f * (1.0f - c_cos20[m]);
        env = fmaf(w, w, env);
      }
    }
    audio[(size_t)seg.out_off + n] = acc / env;
  }
}

}  // namespace

cudaError_t launch_colstats(const float* x, int ldx, int C, const Seg* segs, int nseg, int max_len, double* stats, int ld_stats, cudaStream_t s) {
  if (nseg <= 0 || C <= 0 || max_len <= 0) return cudaSuccess;
  dim3 grid((C + 127) / 128, (max_len + CS_ROWS - 1) / CS_ROWS, nseg), block(32, 8);
  colstats_kernel<<<grid, block, 0, s>>>(x, ldx, C, segs, stats, ld_stats);
  return launched();
}
cudaError_t launch_zero_f64(double* p, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  zero_f64_kernel<<<(unsigned)min((size_t)1184, (n + 255) / 256), 256, 0, s>>>(p, n);
  return launched();
}
cudaError_t launch_adain_snake(const float* x, int ldx, float* y, int ldy, int C, const double* pstats, int ld_pstats, const float* gb, int ld_gb,
                               int beta_off, const float* alpha, const Seg* segs, int nseg, int max_len, cudaStream_t s) {
  if (nseg <= 0 || max_len <= 0) return cudaSuccess;
  if ((C & 3) || (ldx & 3) || (ldy & 3)) return cudaErrorInvalidValue;
  dim3 grid((max_len + AS_ROWS - 1) / AS_ROWS, nseg);
  adain_snake_kernel<<<grid, 256, 5 * C * sizeof(float), s>>>(x, ldx, y, ldy, C, pstats, ld_pstats, gb, ld_gb, beta_off, alpha, segs);
  return launched();
}
cudaError_t launch_pool_up2(const float* x, int ldx, float* y, int ldy, int C, const float* w, const float* bias,
                            const double* pstats, int ld_pstats, const float* gb, int ld_gb, int beta_off, float slope,
                            const Seg* segs, int nseg, int max_out_len, cudaStream_t s) {
  if (nseg <= 0 || max_out_len <= 0) return cudaSuccess;
  dim3 grid((C + 127) / 128, (max_out_len + 63) / 64, nseg);
  pool_up2_kernel<<<grid, 128, 0, s>>>(x, ldx, y, ldy, C, w, bias, pstats, ld_pstats, gb, ld_gb, beta_off, slope, segs);
  return launched();
}
cudaError_t launch_conv1_s2(const float* x, float* y, int ldy, const float* w3, const float* bias1, const Seg* segs, int nseg, int max_out_len,
                            double* ostats, int ld_ostats, cudaStream_t s) {
  (void)ostats; (void)ld_ostats;
  if (nseg <= 0 || max_out_len <= 0) return cudaSuccess;
  dim3 grid((max_out_len + 127) / 128, nseg);
  conv1_s2_kernel<<<grid, 128, 0, s>>>(x, y, ldy, w3, bias1, segs);
  return launched();
}
cudaError_t launch_copy_cols(const float* src, int lds, float* dst, int ldd, int C, int rows, cudaStream_t s) {
  if (rows <= 0 || C <= 0) return cudaSuccess;
  const size_t n = (size_t)rows * C;
  copy_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, lds, dst, ldd, C, rows);
  return launched();
}
cudaError_t launch_source_stft(const float* f0, const float* noise, const float* lin_w, const float* lin_b, float* har_src, float* har, int ld_har,
                               float* scratch_phase, const Seg* segs, int nseg, int max_L, unsigned long long seed, const unsigned long long* seg_seeds,
                               cudaStream_t s) {
  if (nseg <= 0 || max_L <= 0) return cudaSuccess;
  cudaError_t e = ensure_tables();
  if (e != cudaSuccess) return e;
  float* phase = scratch_phase;
  static const bool serial = getenv("TTSB_PHASE_SERIAL") != nullptr;
  if (serial) phase_scan_serial_kernel<<<(nseg * 9 + 63) / 64, 64, 0, s>>>(f0, phase, segs, nseg);
  else phase_scan_kernel<<<(nseg * 9 + 3) / 4, 128, 0, s>>>(f0, phase, segs, nseg);
  if ((e = launched()) != cudaSuccess) return e;
  const int max_frames = 60 * max_L + 1;
  dim3 grid((max_frames + SF - 1) / SF, nseg);
  source_stft_kernel<<<grid, 128, 0, s>>>(f0, noise, lin_w, lin_b, har_src, har, ld_har, phase, segs, seed, seg_seeds);
  return launched();
}
cudaError_t launch_up_add(const float* up, int ldu, int up_row_shift, int reflect, const float* xs, int ldxs, float* y, int ldy, int C,
                          const Seg* segs, int nseg, int max_out_len, double* ostats, int ld_ostats, cudaStream_t s) {
  if (nseg <= 0 || max_out_len <= 0) return cudaSuccess;
  dim3 grid((C + 127) / 128, (max_out_len + 63) / 64, nseg);
  up_add_kernel<<<grid, 128, 0, s>>>(up, ldu, up_row_shift, reflect, xs, ldxs, y, ldy, C, segs, ostats, ld_ostats);
  return launched();
}
cudaError_t launch_istft(const float* cp, int ld, float* audio, const Seg* segs, int nseg, int max_frames, bool precise, cudaStream_t s) {
  if (nseg <= 0 || max_frames <= 1) return cudaSuccess;
  cudaError_t e = ensure_tables();
  if (e != cudaSuccess) return e;
  if (ld < 24 || (ld & 3) || (reinterpret_cast<uintptr_t>(cp) & 15)) return cudaErrorInvalidValue;   // rows are read as six float4
  const int max_samples = 5 * (max_frames - 1) + 10;       // padded-signal samples that matter
  dim3 grid((max_samples + 5 * IS_FPB - 1) / (5 * IS_FPB), nseg);
  if (precise) istft_kernel<true><<<grid, IS_FPB + 32, 0, s>>>(cp, ld, audio, segs);
  else istft_kernel<false><<<grid, IS_FPB + 32, 0, s>>>(cp, ld, audio, segs);
  return launched();
}

}  // namespace ttsb
