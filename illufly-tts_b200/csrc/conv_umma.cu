// tcgen05 implicit-GEMM 1-D convolution over ragged channels-last rows (sm_100a only).
// Persistent and warp-specialised: one CTA per SM walks a contiguous range of (Cout tile, 128-row tile) work items; producer
// warps build the operand slab of item i+1 while the tensor core works on item i and the epilogue warps drain item i-1
// (two slabs, 3-4 weight stages, two TMEM accumulators).
//
//   D[time (M = 128 TMEM lanes), Cout tile (N <= 256 TMEM columns)] += A[time + tap*dil - pad, Cin chunk] * W[tap][Cin chunk][Cout]
//
// * A operand: the CTA's producer warps read the fp32 activations once (128 + halo rows x up to 128 channels per "slab"),
//   apply the fused prologue (InstanceNorm normalise + AdaIN affine + Snake | LeakyReLU) and write bf16 into shared memory in
//   the no-swizzle K-major UMMA layout [channel-chunk plane][row][8 ch = 16 B].  A tap is then just a row offset of the
//   operand descriptor's start address (row * 16 B) — one slab serves all k taps, no im2col, no re-read.
// * B operand: weights pre-packed on the host into the exact shared-memory image of one (tap, 64-channel) stage, streamed by
//   the bulk-copy engine (cp.async.bulk, UBLKCP) through a 3-stage mbarrier ring.
// * MMA: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 x N x K16, bf16 x bf16 -> fp32 in TMEM);
//   tcgen05.commit releases weight stages / slabs and publishes the accumulator.
// * Epilogue (the 8 producer warps): tcgen05.ld 32x32b.x32 -> per-warp shared staging tile -> coalesced float4 rows:
//   + bias, + residual, * scale, (+ running sum of the three parallel resblocks), and per-(utterance, channel) sum / sum of
//   squares for the NEXT AdaIN, so InstanceNorm never costs a pass of its own.
// Reference maths: SURVEY.md A.7, A.11, A.12 (kokoro AdainResBlk1d / Generator / AdaINResBlock1).
#include "common.cuh"
#include "umma_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

namespace ttsb {

using namespace umma;

namespace {

constexpr int BM = 128;
constexpr int KCH = 64;                 // channels per weight stage
constexpr int NPW = 8;                  // producer / epilogue warps
constexpr int NPROD = NPW * 32;
constexpr int NEPI = 128;               // 4 epilogue warps (one per TMEM lane quadrant)
constexpr int NTHREADS = NPROD + 64 + NEPI + 32;   // producers (0-7), weight loader (8), MMA issuer (9), epilogue (10-13), row loader (14)
constexpr int MAXRAW = 8;               // slots of the raw-row ring
constexpr int MAXSTAGE = 8;
// split-fp16 mode: operands are pre-scaled by powers of two so that the low halves stay in fp16's normal range
// (activations by a fixed 16: |x| > 4094 would saturate and is counted in ConvArgs::sat; weights by a per-layer power of two chosen
// at pack time from max|w|, ConvArgs::split_wscale)
constexpr float SPLIT_SA = 16.f;
// Snake's cosine: the fast intrinsic in the 16-bit operand modes (its 2^-21 absolute error is far below the operand rounding),
// the accurate one in the split-fp16 (fp32-class) mode
template <int WFMT> __device__ __forceinline__ float snake_cos(float t) { return WFMT == 2 ? cosf(t) : __cosf(t); }
constexpr int STG_LD = 36;              // floats per staging row (32 + 4 pad: 16-B aligned rows, conflict-free both ways)
constexpr int STG_BYTES = 32 * STG_LD * 4;

struct Plan {
  int BN, KS, nsc, SR, kcn, tmem_cols, ntile_n, nstage, items, per_cta, split, acc_cols;
  int nraw, rch;                    // raw-row ring: slots and rows per slot (0 slots = ring disabled)
  uint32_t off_bar, off_stats, off_par, off_stg, off_slab, slab_bytes, off_w, wstage_bytes, smem_bytes, off_raw, raw_bytes;
};

struct UArgs {
  ConvArgs a;
  Plan p;
  long long* prof;   // optional [grid][16] cycle counters per role (TTSB_UMMA_PROF), else null
  const uint8_t* xpack;   // pre-packed operand slabs [tile][K chunk][slab image] (pack_split_kernel), else null
  int l2pf;    // row loader asks the L2 for the next item's inputs (TTSB_L2PF bit 0; default off: measured slower on these layers)
  int dbg;     // TTSB_UMMA_DBG (timing experiments only, results become garbage): 1 = producers skip global loads,
               // 2 = epilogue skips global traffic, 4 = no MMA issue; applied to the Snake-prologue variants only
};

// One 32-row x 32-column chunk of the epilogue for a warp, fast path: the tile is fully inside the utterance, the channel
// group is complete, no activation.  y = (acc * acs + bias [+ res]) [* os] [+ y_old]; returns the lane's partial column sums.
// RES / SCALE / ACC are compile-time so that each launch runs one short loop (4 epilogue warps are issue-latency bound:
// ncu shows ~8 cycles per instruction per warp, so the instruction count per element is what matters).
// r4: the residual rows of this chunk, requested by the caller at the TOP of the chunk - before the TMEM load and the staging
// writes, so that their latency overlaps that work (the L2 prefetch of the row loader has made them L2 hits by then).
template <bool RES, bool SCALE, bool ACC>
__device__ __forceinline__ void epi_chunk_fast(const float* __restrict__ stg, int lane, float acs, float os, const float4 b4, float* __restrict__ yb,
                                               size_t ystep, const float4 (&r4)[8], float (&ps)[4], float (&pq)[4]) {
  float4 y4[8];
  if (ACC) {
#pragma unroll
    for (int i = 0; i < 8; ++i) y4[i] = *reinterpret_cast<const float4*>(yb + i * ystep);
  }
  const float* sp = stg + (lane >> 3) * STG_LD + (lane & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 a4 = *reinterpret_cast<const float4*>(sp + i * 4 * STG_LD);
    float4 o;
    o.x = fmaf(a4.x, acs, b4.x); o.y = fmaf(a4.y, acs, b4.y); o.z = fmaf(a4.z, acs, b4.z); o.w = fmaf(a4.w, acs, b4.w);
    if (RES) { o.x += r4[i].x; o.y += r4[i].y; o.z += r4[i].z; o.w += r4[i].w; }
    if (SCALE && ACC) { o.x = fmaf(o.x, os, y4[i].x); o.y = fmaf(o.y, os, y4[i].y); o.z = fmaf(o.z, os, y4[i].z); o.w = fmaf(o.w, os, y4[i].w); }
    else if (SCALE) { o.x *= os; o.y *= os; o.z *= os; o.w *= os; }
    else if (ACC) { o.x += y4[i].x; o.y += y4[i].y; o.z += y4[i].z; o.w += y4[i].w; }
    *reinterpret_cast<float4*>(yb + i * ystep) = o;
    ps[0] += o.x; ps[1] += o.y; ps[2] += o.z; ps[3] += o.w;
    pq[0] = fmaf(o.x, o.x, pq[0]); pq[1] = fmaf(o.y, o.y, pq[1]); pq[2] = fmaf(o.z, o.z, pq[2]); pq[3] = fmaf(o.w, o.w, pq[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// PRO (fused operand prologue) and WFMT (operand format) are compile-time: the producers' per-element code is the
// instruction-issue hot spot of the kernel (ncu), so no run-time switches inside it.
template <int PRO, int WFMT>
__global__ void __launch_bounds__(NTHREADS, 1) conv_umma_kernel(const UArgs u) {
  extern __shared__ __align__(128) uint8_t smem[];
  const ConvArgs& p = u.a;
  const Plan& pl = u.p;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = pl.nstage;

  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem);
  const uint32_t bar0 = smem_u32(smem + pl.off_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (MAXSTAGE + s); };
  auto bar_slab_full = [&](int b) { return bar0 + 8u * (2 * MAXSTAGE + b); };
  auto bar_slab_empty = [&](int b) { return bar0 + 8u * (2 * MAXSTAGE + 2 + b); };
  auto bar_acc_full = [&](int b) { return bar0 + 8u * (2 * MAXSTAGE + 4 + b); };
  auto bar_acc_empty = [&](int b) { return bar0 + 8u * (2 * MAXSTAGE + 6 + b); };
  auto bar_raw_full = [&](int b) { return bar0 + 8u * (2 * MAXSTAGE + 8 + b); };
  auto bar_raw_empty = [&](int b) { return bar0 + 8u * (2 * MAXSTAGE + 8 + MAXRAW + b); };
  double* s_sum = reinterpret_cast<double*>(smem + pl.off_stats);   // [4 row quadrants][256], double: the running sums
  double* s_sq = s_sum + 4 * 256;                                   // must not depend on how tiles are dealt to CTAs
  float* s_par = reinterpret_cast<float*>(smem + pl.off_par);       // [5][cin_pad]: mean, a, b, alpha, 1/alpha
  const int cpad = p.cin_pad;

  // this CTA's contiguous range of work items; item -> (Cout tile nt, row tile)
  const int item0 = blockIdx.x * pl.per_cta;
  const int item1 = min(item0 + pl.per_cta, pl.items);

  // ---- one-time setup ----
  if (warp == NPW + 1) tmem_alloc(smem_u32(s_tmem), (uint32_t)pl.tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < MAXSTAGE; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      // one elected lane arrives for its warp (after __syncwarp): 256 lanes arriving on one shared-memory word serialise
      mbar_init(bar_slab_full(b), u.xpack ? 1 : NPROD / 32); mbar_init(bar_slab_empty(b), 1);
      mbar_init(bar_acc_full(b), 1); mbar_init(bar_acc_empty(b), NEPI / 32);
    }
    for (int b = 0; b < MAXRAW; ++b) { mbar_init(bar_raw_full(b), 1); mbar_init(bar_raw_empty(b), NPROD / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int c = tid; c < 2048; c += NTHREADS) s_sum[c] = 0.0;
  pdl_trigger();                                   // the next kernel on the stream may start its own prologue
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // everything above ran under the previous kernel's tail; the weight loader (static data) keeps going, every other role waits here
  // for the previous kernel's results (activations, statistics, residuals, packed operand images)
  if (warp != NPW) pdl_wait();

  const int nwst = pl.KS / KCH;                      // weight stages per (slab, tap)
  const uint32_t slab0 = smem_u32(smem + pl.off_slab);
  const uint32_t w0 = smem_u32(smem + pl.off_w);

  if (warp < NPW && u.xpack) {
    // =================== pre-packed operands: the slabs are bulk-copied as they lie in global memory ===================
    if (tid == 0) {
      uint32_t g = 0;
      long long t_wait = 0;
      const long long t_start = clock64();
      for (int item = item0; item < item1; ++item) {
        const uint8_t* src = u.xpack + (size_t)(item % p.ntiles) * pl.nsc * pl.slab_bytes;
        for (int sc = 0; sc < pl.nsc; ++sc, ++g) {
          const uint32_t b = g & 1u;
          mbar_wait_t(bar_slab_empty(b), ((g >> 1) & 1u) ^ 1u, t_wait);
          mbar_expect_tx(bar_slab_full(b), pl.slab_bytes);
          bulk_g2s(slab0 + b * pl.slab_bytes, src + (size_t)sc * pl.slab_bytes, pl.slab_bytes, bar_slab_full(b));
        }
      }
      if (u.prof) { u.prof[blockIdx.x * 16 + 0] = clock64() - t_start; u.prof[blockIdx.x * 16 + 1] = t_wait; }
    }
  } else if (warp < NPW) {
    // =================== producers: build the 16-bit operand slabs ===================
    const int planes = pl.KS >> 3;
    const int rpp = NPROD / planes;                    // rows per pass
    const int pln = tid % planes, r0 = tid / planes;
    int cur_seg = -1, par_c0 = -1, par_seg = -1;
    float pA[8], pB[8], pA2[8], pB2[8], pK[8];
    float satm = 0.f;                                  // split mode: largest |operand| this thread converted (range check, ConvArgs::sat)
    uint32_t g = 0;                                    // running slab counter of this CTA
    uint32_t rslot = 0, rphase = 0;                    // position in the raw-row ring
    long long t_wait = 0;
    const long long t_start = clock64();
    for (int item = item0; item < item1; ++item) {
      const Tile tile = p.tiles[item % p.ntiles];
      const Seg seg = p.segs[tile.seg];
      if (PRO >= PRO_ADAIN_LRELU && tile.seg != cur_seg) {
        // per-(utterance, channel) AdaIN coefficients; only producers read them, so a producer-only barrier suffices
        asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
        for (int c = tid; c < cpad; c += NPROD) {
          float mean = 0.f, a = 0.f, b = 0.f, al = 1.f;
          if (c < p.Cin) {
            const double* st = p.pstats + ((size_t)tile.seg * p.ld_pstats + c) * 2;
            adain_coef(st, seg.in_len, p.gb[(size_t)tile.seg * p.ld_gb + c], p.gb[(size_t)tile.seg * p.ld_gb + p.beta_off + c], mean, a, b);
            if (PRO == PRO_ADAIN_SNAKE) al = p.alpha[c];
          }
          // folded constants: AdaIN is one FFMA  w = x A + B  (A = (1+gamma) rstd, B = beta - mean A);
          // Snake  w + sin^2(alpha w)/alpha = (w + k) - k cos(2 alpha w),  k = 1/(2 alpha):  wk = x A + (B + k),  t = x (2 alpha A) + 2 alpha B
          const float Bf = fmaf(-mean, a, b);
          const float kk = 0.5f / al;
          s_par[c] = a; s_par[cpad + c] = (PRO == PRO_ADAIN_SNAKE) ? Bf + kk : Bf;
          s_par[2 * cpad + c] = 2.f * al * a; s_par[3 * cpad + c] = 2.f * al * Bf; s_par[4 * cpad + c] = kk;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
        cur_seg = tile.seg;
      }
      for (int sc = 0; sc < pl.nsc; ++sc, ++g) {
        const uint32_t b = g & 1u;
        mbar_wait_t(bar_slab_empty(b), ((g >> 1) & 1u) ^ 1u, t_wait);
        const int c0 = sc * pl.KS + pln * 8;
        if (PRO >= PRO_ADAIN_LRELU && (c0 != par_c0 || tile.seg != par_seg)) {      // one slab per tile: the constants stay in registers
          par_c0 = c0; par_seg = tile.seg;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            pA[e] = s_par[c0 + e]; pB[e] = s_par[cpad + c0 + e];
            if (PRO == PRO_ADAIN_SNAKE) { pA2[e] = s_par[2 * cpad + c0 + e]; pB2[e] = s_par[3 * cpad + c0 + e]; pK[e] = s_par[4 * cpad + c0 + e]; }
          }
        }
        const bool tailc = c0 + 8 > p.Cin;                 // only the last channel chunk needs per-element masking
        uint8_t* slab = smem + pl.off_slab + (size_t)b * pl.slab_bytes + (size_t)pln * pl.SR * 16;
        const size_t lo_off = (size_t)planes * pl.SR * 16;      // split mode: the low halves follow the high planes
        const bool second = c0 + 4 < p.ldx;
        // software-pipelined batches of 4 rows: the eight 16-B loads of batch n+1 are in flight while batch n is transformed
        // (ncu: the producer was stalled on long_scoreboard right after its loads; 1 CTA/SM leaves no other warps to hide it)
        constexpr int NB = 2;        // rows per batch (two batches live in registers)
        auto load_batch = [&](int rbase, float (&v)[NB][8], bool (&ok)[NB]) {
#pragma unroll
          for (int h = 0; h < NB; ++h) {
            const int rr = rbase + h * rpp;
            const int traw = tile.t0 + rr - p.pad;
            const int tin = traw >> p.in_shr;
            ok[h] = rr < pl.SR && traw >= 0 && tin < seg.in_len && c0 < p.Cin && !(u.dbg & 1);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[h][e] = 0.f;
            if (ok[h]) {
              const float* src = p.x + (size_t)(seg.in_off + tin) * p.ldx + c0;
              const float4 q0 = __ldg(reinterpret_cast<const float4*>(src));
              v[h][0] = q0.x; v[h][1] = q0.y; v[h][2] = q0.z; v[h][3] = q0.w;
              if (second) {
                const float4 q1 = __ldg(reinterpret_cast<const float4*>(src + 4));
                v[h][4] = q1.x; v[h][5] = q1.y; v[h][6] = q1.z; v[h][7] = q1.w;
              }
            }
          }
        };
        auto finish_batch = [&](int rbase, float (&v)[NB][8], const bool (&ok)[NB]) {
#pragma unroll
          for (int h = 0; h < NB; ++h) {
            const int rr = rbase + h * rpp;
            if (rr >= pl.SR) break;
            if (ok[h]) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float x = v[h][e];
                if (PRO == PRO_LRELU) x = fmaxf(x, x * p.slope);                    // slope < 1
                else if (PRO == PRO_ADAIN_LRELU) { const float wv = fmaf(x, pA[e], pB[e]); x = fmaxf(wv, wv * p.slope); }
                else if (PRO == PRO_ADAIN_SNAKE) x = fmaf(-pK[e], snake_cos<WFMT>(fmaf(x, pA2[e], pB2[e])), fmaf(x, pA[e], pB[e]));
                v[h][e] = x;
              }
              if (tailc) {
#pragma unroll
                for (int e = 0; e < 8; ++e) if (c0 + e >= p.Cin) v[h][e] = 0.f;
              }
            }
            uint4 o;
            if (WFMT == 2) {
              uint4 ol;
              float hi[8], lo[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                satm = fmaxf(satm, fabsf(v[h][e]));
                const float xs = fminf(fmaxf(v[h][e] * SPLIT_SA, -65504.f), 65504.f);
                hi[e] = __half2float(__float2half_rn(xs));
                lo[e] = xs - hi[e];                        // exact in fp32; its own fp16 rounding error is ~2^-22 of x
              }
              o.x = pack_f16(hi[0], hi[1]); o.y = pack_f16(hi[2], hi[3]); o.z = pack_f16(hi[4], hi[5]); o.w = pack_f16(hi[6], hi[7]);
              ol.x = pack_f16(lo[0], lo[1]); ol.y = pack_f16(lo[2], lo[3]); ol.z = pack_f16(lo[4], lo[5]); ol.w = pack_f16(lo[6], lo[7]);
              *reinterpret_cast<uint4*>(slab + lo_off + (size_t)rr * 16) = ol;
            }
            else if (WFMT == 1) {
              if (PRO >= PRO_ADAIN_LRELU) {      // instance-normalised data is bounded (|x^| <= sqrt(n)): no saturation needed
                o.x = pack_f16_nosat(v[h][0], v[h][1]); o.y = pack_f16_nosat(v[h][2], v[h][3]); o.z = pack_f16_nosat(v[h][4], v[h][5]); o.w = pack_f16_nosat(v[h][6], v[h][7]);
              } else {
                o.x = pack_f16(v[h][0], v[h][1]); o.y = pack_f16(v[h][2], v[h][3]); o.z = pack_f16(v[h][4], v[h][5]); o.w = pack_f16(v[h][6], v[h][7]);
              }
            }
            else { o.x = pack_bf16(v[h][0], v[h][1]); o.y = pack_bf16(v[h][2], v[h][3]); o.z = pack_bf16(v[h][4], v[h][5]); o.w = pack_bf16(v[h][6], v[h][7]); }
            *reinterpret_cast<uint4*>(slab + (size_t)rr * 16) = o;
          }
        };
        // interior tiles (every slab row inside the utterance, full 8-channel chunk): straight-line code, no per-row predicates,
        // four rows (eight 16-B loads) in flight per thread.  Edge tiles of an utterance take the generic path below.
        const int row_lo = tile.t0 - p.pad;
        const bool interior = p.in_shr == 0 && row_lo >= 0 && row_lo + pl.SR <= seg.in_len && c0 + 8 <= p.Cin && second;
        auto xform_store = [&](int rr, const float4 q0, const float4 q1) {
          float v[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float x = v[e];
            if (PRO == PRO_LRELU) x = fmaxf(x, x * p.slope);
            else if (PRO == PRO_ADAIN_LRELU) { const float wv = fmaf(x, pA[e], pB[e]); x = fmaxf(wv, wv * p.slope); }
            else if (PRO == PRO_ADAIN_SNAKE) x = fmaf(-pK[e], snake_cos<WFMT>(fmaf(x, pA2[e], pB2[e])), fmaf(x, pA[e], pB[e]));
            v[e] = x;
          }
          uint4 o;
          if (WFMT == 2) {
            uint4 ol;
            float hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              satm = fmaxf(satm, fabsf(v[e]));
              const float xs = fminf(fmaxf(v[e] * SPLIT_SA, -65504.f), 65504.f);
              hi[e] = __half2float(__float2half_rn(xs));
              lo[e] = xs - hi[e];
            }
            o.x = pack_f16_nosat(hi[0], hi[1]); o.y = pack_f16_nosat(hi[2], hi[3]); o.z = pack_f16_nosat(hi[4], hi[5]); o.w = pack_f16_nosat(hi[6], hi[7]);
            ol.x = pack_f16_nosat(lo[0], lo[1]); ol.y = pack_f16_nosat(lo[2], lo[3]); ol.z = pack_f16_nosat(lo[4], lo[5]); ol.w = pack_f16_nosat(lo[6], lo[7]);
            *reinterpret_cast<uint4*>(slab + lo_off + (size_t)rr * 16) = ol;
          } else if (WFMT == 1) {
            if (PRO >= PRO_ADAIN_LRELU) { o.x = pack_f16_nosat(v[0], v[1]); o.y = pack_f16_nosat(v[2], v[3]); o.z = pack_f16_nosat(v[4], v[5]); o.w = pack_f16_nosat(v[6], v[7]); }
            else { o.x = pack_f16(v[0], v[1]); o.y = pack_f16(v[2], v[3]); o.z = pack_f16(v[4], v[5]); o.w = pack_f16(v[6], v[7]); }
          } else { o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]); }
          *reinterpret_cast<uint4*>(slab + (size_t)rr * 16) = o;
        };
        const bool ring = pl.nraw > 0 && p.in_shr == 0 && row_lo >= 0 && row_lo + pl.SR <= seg.in_len;
        if (ring) {
          // interior tile: the row loader streams the fp32 rows of this slab through a shared-memory ring with the bulk-copy
          // engine (no LDG, no L1 queue); the producers only transform ring slot -> 16-bit slab, two rows in flight per thread
          // (a slot holds 2 * rpp rows)
          const int nch = (pl.SR + pl.rch - 1) / pl.rch;
          // groups of four lanes alternate which 16-B half of their 32 B they read first: eight lanes x 16 B at a 32-B stride
          // would hit every bank group twice
          const int hf = (tid >> 2) & 1;
          const int sp_off = r0 * pl.KS + pln * 8;
          const int row2 = rpp * pl.KS;                     // floats between a thread's two rows of a slot
          for (int ch = 0; ch < nch; ++ch) {
            mbar_wait_t(bar_raw_full(rslot), rphase, t_wait);
            const float* sp = reinterpret_cast<const float*>(smem + pl.off_raw + (size_t)rslot * pl.raw_bytes) + sp_off;
            const int ra = ch * pl.rch + r0, rb = ra + rpp;
            const bool hasa = ra < pl.SR, hasb = rb < pl.SR;
            const float* spb = hasb ? sp + row2 : sp;
            float4 ax = *reinterpret_cast<const float4*>(sp + 4 * hf);
            float4 ay = *reinterpret_cast<const float4*>(sp + 4 * (hf ^ 1));
            float4 bx = *reinterpret_cast<const float4*>(spb + 4 * hf);
            float4 by = *reinterpret_cast<const float4*>(spb + 4 * (hf ^ 1));
            float4 a0 = hf ? ay : ax, a1 = hf ? ax : ay, b0 = hf ? by : bx, b1 = hf ? bx : by;
            if (tailc) {        // channels past Cin: the ring holds stale bytes there and the operand must be zero
              float* t0 = &a0.x; float* t1 = &a1.x; float* t2 = &b0.x; float* t3 = &b1.x;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (c0 + e >= p.Cin) { t0[e] = 0.f; t2[e] = 0.f; }
                if (c0 + 4 + e >= p.Cin) { t1[e] = 0.f; t3[e] = 0.f; }
              }
            }
            if (hasa) xform_store(ra, a0, a1);
            if (hasb) xform_store(rb, b0, b1);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_raw_empty(rslot));      // this warp is done reading the slot
            if (++rslot == (uint32_t)pl.nraw) { rslot = 0; rphase ^= 1u; }
          }
        } else if (c0 >= p.Cin) {                      // padding channels of the last K chunk: zeros, no loads
          const uint4 z = make_uint4(0u, 0u, 0u, 0u);
          for (int r = r0; r < pl.SR; r += rpp) {
            *reinterpret_cast<uint4*>(slab + (size_t)r * 16) = z;
            if (WFMT == 2) *reinterpret_cast<uint4*>(slab + lo_off + (size_t)r * 16) = z;
          }
        } else if (interior) {
          const float* src0 = p.x + (size_t)(seg.in_off + row_lo) * p.ldx + c0;
          int r = r0;
          for (; r + 3 * rpp < pl.SR; r += 4 * rpp) {
            float4 qa[4], qb[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const float* sp = src0 + (size_t)(r + h * rpp) * p.ldx;
              qa[h] = __ldg(reinterpret_cast<const float4*>(sp));
              qb[h] = __ldg(reinterpret_cast<const float4*>(sp + 4));
            }
#pragma unroll
            for (int h = 0; h < 4; ++h) xform_store(r + h * rpp, qa[h], qb[h]);
          }
          for (; r < pl.SR; r += rpp) {
            const float* sp = src0 + (size_t)r * p.ldx;
            xform_store(r, __ldg(reinterpret_cast<const float4*>(sp)), __ldg(reinterpret_cast<const float4*>(sp + 4)));
          }
        } else {
          float va[NB][8];
          bool oka[NB];
#pragma unroll 1
          for (int r = r0; r < pl.SR; r += NB * rpp) {
            load_batch(r, va, oka);
            finish_batch(r, va, oka);
          }
        }
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_slab_full(b));
      }
    }
    if (WFMT == 2 && p.sat && satm * SPLIT_SA > 65504.f) atomicAdd(p.sat, 1u);
    if (u.prof && tid == 0) { u.prof[blockIdx.x * 16 + 0] = clock64() - t_start; u.prof[blockIdx.x * 16 + 1] = t_wait; }
  } else if (warp == NPW) {
    // =================== weight loader (bulk-copy engine) ===================
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = item0; item < item1; ++item) {
        const int nt = item / p.ntiles;
        const size_t stage_elems = (size_t)pl.wstage_bytes / 2;
        const uint16_t* wsrc = p.wb + (size_t)nt * p.k * pl.kcn * stage_elems;
        for (int sc = 0; sc < pl.nsc; ++sc)
          for (int j = 0; j < p.k; ++j)
            for (int kc = 0; kc < nwst; ++kc, ++it) {
              const uint32_t s = it % NS;
              mbar_wait(bar_empty(s), ((it / NS) & 1u) ^ 1u);
              mbar_expect_tx(bar_full(s), pl.wstage_bytes);
              const size_t blk = (size_t)j * pl.kcn + (size_t)sc * nwst + kc;
              bulk_g2s(w0 + s * pl.wstage_bytes, wsrc + blk * stage_elems, pl.wstage_bytes, bar_full(s));
            }
      }
    }
  } else if (warp == NPW + 1) {
    // =================== MMA issuer ===================
    // The whole warp walks the loops with warp-uniform values (descriptors then live in uniform registers) and one elected
    // lane issues tcgen05.mma / tcgen05.commit.  Issuing from inside `if (lane == 0)` made the compiler shuttle every operand
    // through R2UR and wrap each UTCHMMA in an ELECT loop: ~190 cycles per MMA against 64-128 cycles of tensor-core time.
    {
      uint32_t leader;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t ofmt = WFMT >= 1 ? 0u : 1u;       // instruction-descriptor operand format: 0 = F16, 1 = BF16
      const uint32_t a_lo = (uint32_t)(pl.KS >> 3) * (uint32_t)pl.SR * 16u;     // split mode: offset of the low-half planes
      const uint32_t b_lo = (uint32_t)pl.BN * KCH * 2u;                         //             and of the low-half weight block
      const uint32_t idesc = (1u << 4) | (ofmt << 7) | (ofmt << 10) | ((uint32_t)(pl.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t a_lbo = (uint32_t)pl.SR * 16u, a_sbo = 128u;
      const uint32_t b_lbo = (uint32_t)pl.BN * 16u, b_sbo = 128u;
      const uint64_t a_desc0 = make_desc(0, a_lbo, a_sbo), b_desc0 = make_desc(0, b_lbo, b_sbo);
      const uint32_t a_kstep = (uint32_t)(2 * pl.SR);          // two 16-B channel planes per K=16 step, in 16-B units
      const uint32_t b_kstep = (uint32_t)(2 * pl.BN);
      uint32_t it = 0, g = 0, n = 0;
      long long tw_slab = 0, tw_acc = 0, tw_w = 0;
      const long long t_start = clock64();
      for (int item = item0; item < item1; ++item, ++n) {
        const uint32_t slot = n & 1u;
        mbar_wait_t(bar_acc_empty(slot), ((n >> 1) & 1u) ^ 1u, tw_acc);       // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_u + slot * (uint32_t)pl.acc_cols;
        const uint32_t tcor = tacc + (uint32_t)pl.BN;       // split mode: the small cross terms get their own accumulator
        uint32_t first = 1;
        for (int sc = 0; sc < pl.nsc; ++sc, ++g) {
          const uint32_t b = g & 1u;
          mbar_wait_t(bar_slab_full(b), (g >> 1) & 1u, tw_slab);
          tc_fence_after();
          const uint32_t sbase = slab0 + b * pl.slab_bytes;
          for (int j = 0; j < p.k; ++j)
            for (int kc = 0; kc < nwst; ++kc, ++it) {
              const uint32_t s = it % NS;
              mbar_wait_t(bar_full(s), (it / NS) & 1u, tw_w);
              tc_fence_after();
              const uint32_t wbase = w0 + s * pl.wstage_bytes;
              // descriptors differ only in the 14-bit start-address field (address >> 4): one 64-bit add per operand per MMA.
              // (The issuing thread's own instruction latency bounds the MMA rate: measured 960 cycles per 4-MMA stage with
              //  the descriptors rebuilt from scratch, against 256 cycles of tensor-core time.)
              const uint64_t ad0 = a_desc0 + (uint64_t)((sbase >> 4) + (uint32_t)(kc * 8 * pl.SR + j * p.dil));
              const uint64_t bd0 = b_desc0 + (uint64_t)(wbase >> 4);
#pragma unroll
              for (int kk = 0; kk < KCH / 16; ++kk) {
                const uint64_t ad = ad0 + (uint64_t)(kk * a_kstep);
                const uint64_t bd = bd0 + (uint64_t)(kk * b_kstep);
                if (leader && !(u.dbg & 4)) umma_bf16(tacc, ad, bd, idesc, first ? 0u : 1u);
                if (WFMT == 2) {
                  // x w ~= xh wh + (xh wl + xl wh)   (xl wl is ~2^-22 relative and dropped).  The tensor core truncates at
                  // every accumulate, an error that grows linearly with the number of MMAs into one accumulator: the cross
                  // terms (2^-11 of the main term) go to their own accumulator, so the main one sees K/16 steps, not 3K/16
                  if (leader) {
                    umma_bf16(tcor, ad, bd + (uint64_t)(b_lo >> 4), idesc, first ? 0u : 1u);
                    umma_bf16(tcor, ad + (uint64_t)(a_lo >> 4), bd, idesc, 1u);
                  }
                }
                first = 0;
              }
              if (leader) umma_commit(bar_empty(s));            // weight stage free once these MMAs have read it
              __syncwarp();
            }
          if (leader) umma_commit(bar_slab_empty(b));           // slab free once every MMA issued so far has completed
        }
        if (leader) umma_commit(bar_acc_full(slot));            // accumulator of this item complete
      }
      if (u.prof && lane == 0) {
        long long* pp = u.prof + blockIdx.x * 16;
        pp[2] = clock64() - t_start; pp[3] = tw_slab; pp[4] = tw_acc; pp[5] = tw_w;
      }
    }
  } else if (warp == NPW + 2 + NEPI / 32) {
    // =================== row loader: L2 prefetch one item ahead; fp32 activation rows -> raw ring (bulk-copy engine) ==========
    {
      uint32_t rslot = 0, rphase = 0;
      const int nch = (pl.SR + pl.rch - 1) / pl.rch;
      // One item ahead, ask the L2 for what the producers (activation rows or packed operand images) and the epilogue (residual
      // rows, the old output when it is accumulated) will read: their requests then find the lines in L2 instead of paying the
      // HBM latency with the few KB each role can keep in flight.
      auto prefetch_item = [&](int item) {
        if (item >= item1 || lane != 0 || !u.l2pf) return;
        const Tile tl = p.tiles[item % p.ntiles];
        const Seg sg = p.segs[tl.seg];
        auto pf = [&](const void* ptr, size_t bytes) {
          const uint8_t* b = reinterpret_cast<const uint8_t*>(ptr);
          const size_t lead = (16 - (reinterpret_cast<uintptr_t>(b) & 15)) & 15;      // 16-byte aligned, a multiple of 16
          if (bytes <= lead) return;
          b += lead; bytes = (bytes - lead) & ~size_t(15);
          for (size_t o = 0; o < bytes; o += 16384) prefetch_l2(b + o, (uint32_t)(bytes - o < 16384 ? bytes - o : 16384));
        };
        if (!(u.l2pf & 1)) {
        } else if (u.xpack) {
          pf(u.xpack + (size_t)(item % p.ntiles) * pl.nsc * pl.slab_bytes, (size_t)pl.nsc * pl.slab_bytes);
        } else {
          int lo = (tl.t0 - p.pad) >> p.in_shr, hi = ((tl.t0 - p.pad + pl.SR - 1) >> p.in_shr) + 1;
          lo = max(lo, 0); hi = min(hi, sg.in_len);
          if (hi > lo) pf(p.x + (size_t)(sg.in_off + lo) * p.ldx, (size_t)(hi - lo) * p.ldx * 4);
        }
        const int nr = min(BM, sg.out_len - tl.t0);
        const int n0 = (item / p.ntiles) * pl.BN;
        if (!(u.l2pf & 4)) return;
        if (nr > 0 && p.res) {
          if (p.ldr <= pl.BN) pf(p.res + (size_t)(sg.out_off + tl.t0) * p.ldr, (size_t)nr * p.ldr * 4);      // whole rows
          else for (int r = 0; r < nr; ++r) pf(p.res + (size_t)(sg.out_off + tl.t0 + r) * p.ldr + n0, (size_t)min(pl.BN, p.Cout - n0) * 4);
        }
        if (nr > 0 && p.accumulate && p.ldy <= pl.BN) pf(p.y + (size_t)(sg.out_off + tl.t0) * p.ldy, (size_t)nr * p.ldy * 4);
      };
      prefetch_item(item0);
      for (int item = item0; item < item1; ++item) {
        prefetch_item(item + 1);
        if (pl.nraw == 0) continue;
        const Tile tile = p.tiles[item % p.ntiles];
        const Seg seg = p.segs[tile.seg];
        const int row_lo = tile.t0 - p.pad;
        if (!(p.in_shr == 0 && row_lo >= 0 && row_lo + pl.SR <= seg.in_len)) continue;      // edge tile: producers load it themselves
        for (int sc = 0; sc < pl.nsc; ++sc) {
          const int cbase = sc * pl.KS;
          int nfl = p.ldx - cbase;                        // floats of this channel slice present in a row of x
          nfl = nfl > pl.KS ? pl.KS : nfl;
          const uint32_t row_bytes = (uint32_t)nfl * 4u;
          const bool contiguous = p.ldx == pl.KS;         // whole rows: the chunk is one contiguous block
          for (int ch = 0; ch < nch; ++ch) {
            const uint32_t slot = rslot;
            mbar_wait(bar_raw_empty(slot), rphase ^ 1u);
            if (++rslot == (uint32_t)pl.nraw) { rslot = 0; rphase ^= 1u; }
            const int rows_n = min(pl.rch, pl.SR - ch * pl.rch);
            const float* src = p.x + (size_t)(seg.in_off + row_lo + ch * pl.rch) * p.ldx + cbase;
            const uint32_t dst = smem_u32(smem + pl.off_raw + (size_t)slot * pl.raw_bytes);
            if (lane == 0) mbar_expect_tx(bar_raw_full(slot), (uint32_t)rows_n * row_bytes);
            __syncwarp();                                  // the byte count is registered before any copy can complete
            if (contiguous) {
              if (lane == 0) bulk_g2s(dst, src, (uint32_t)rows_n * row_bytes, bar_raw_full(slot));
            } else {                                       // strided rows: one copy per row, the lanes issue them side by side
              for (int r = lane; r < rows_n; r += 32)
                bulk_g2s(dst + (uint32_t)r * (uint32_t)pl.KS * 4u, src + (size_t)r * p.ldx, row_bytes, bar_raw_full(slot));
            }
          }
        }
      }
    }
  } else {
    // =================== epilogue (4 warps, one TMEM lane quadrant each) ===================
    const int q = warp & 3;
    const int et = tid - (NPROD + 64);                 // 0..127
    float* stg = reinterpret_cast<float*>(smem + pl.off_stg + (size_t)(warp - (NPW + 2)) * STG_BYTES);
    const int nchunk = pl.BN >> 5;
    const int cc = (lane & 7) * 4;
    const float acs = p.acc_scale != 0.f ? p.acc_scale : 1.0f;
    uint32_t n = 0;
    long long tw_full = 0;
    long long tp_ld = 0, tp_stage = 0, tp_io = 0, tp_stats = 0, tp_pre = 0, tp_top = 0, tp_post = 0;       // TTSB_UMMA_PROF: where an epilogue chunk's cycles go (warp 0 of the role)
#ifdef TTSB_EPI_PROF
    const bool tprof = u.prof != nullptr;
#else
    constexpr bool tprof = false;   // the per-phase timers cost registers: build with -DTTSB_EPI_PROF to get them (profiles/r02_epilogue_phases.md)
#endif
    const long long t_start = clock64();
    for (int item = item0; item < item1; ++item, ++n) {
      long long ti0 = 0;
      if (tprof) ti0 = clock64();
      const int nt = item / p.ntiles;
      const Tile tile = p.tiles[item % p.ntiles];
      const Seg seg = p.segs[tile.seg];
      const int n0 = nt * pl.BN;
      const uint32_t slot = n & 1u;
      const bool tile_fast = tile.t0 + BM <= seg.out_len && p.act == ACT_NONE && !(u.dbg & 2);
      const size_t row0 = (size_t)(seg.out_off + tile.t0 + q * 32 + (lane >> 3));
      const size_t ystep = (size_t)4 * p.ldy, rstep = (size_t)4 * p.ldr;
      // bias of the chunk: requested one chunk ahead (the first before the wait for the accumulator).  Loaded at its point of use it
      // cost a full L2 round trip per chunk: ncu put 23 % of the epilogue warps' stall samples on the first FFMA that reads it.
      float4 b4n = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias && n0 + cc + 3 < p.Cout) b4n = *reinterpret_cast<const float4*>(p.bias + n0 + cc);
      // residual rows: those of the first chunk are requested before the wait for the accumulator, those of chunk c+1 right after
      // chunk c's stores - ahead of its statistics, the next TMEM load and the staging writes, none of which need these registers.
      // Requested at the top of their own chunk they arrived ~900 cycles late (profile: +7.5 k cycles per item on every c2 launch).
      float4 r4[8];
      if (p.res && tile_fast && n0 + cc + 3 < p.Cout) {
        const float* rb = p.res + row0 * p.ldr + n0 + cc;
#pragma unroll
        for (int i = 0; i < 8; ++i) r4[i] = *reinterpret_cast<const float4*>(rb + i * rstep);
      }
      if (tprof) tp_pre += clock64() - ti0;
      mbar_wait_t(bar_acc_full(slot), (n >> 1) & 1u, tw_full);
      tc_fence_after();
      const uint32_t tacc = tmem_base + slot * (uint32_t)pl.acc_cols + ((uint32_t)(q * 32) << 16);
      for (int chunk = 0; chunk < nchunk; ++chunk) {
        long long tc0 = 0;
        if (tprof) tc0 = clock64();
        const int c0 = chunk * 32;
        const int co = n0 + c0 + cc;
        const float4 b4c = b4n;
        b4n = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && chunk + 1 < nchunk && co + 32 + 3 < p.Cout) b4n = *reinterpret_cast<const float4*>(p.bias + co + 32);
        uint32_t v[32];
        long long tq0 = 0;
        if (tprof) { tq0 = clock64(); tp_top += tq0 - tc0; }
        tmem_ld32(tacc + (uint32_t)c0, v);
        if (WFMT == 2) {
          uint32_t v2[32];
          tmem_ld32(tacc + (uint32_t)(pl.BN + c0), v2);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
        }
        if (chunk == nchunk - 1) {                     // last read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty(slot));      // one arrival per warp (128 lanes on one word serialise)
        }
        long long tq1 = 0;
        if (tprof) { tq1 = clock64(); tp_ld += tq1 - tq0; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 f;
          f.x = __uint_as_float(v[4 * i]); f.y = __uint_as_float(v[4 * i + 1]); f.z = __uint_as_float(v[4 * i + 2]); f.w = __uint_as_float(v[4 * i + 3]);
          *reinterpret_cast<float4*>(stg + lane * STG_LD + 4 * i) = f;
        }
        __syncwarp();
        long long tq2 = 0;
        if (tprof) { tq2 = clock64(); tp_stage += tq2 - tq1; }
        float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
        const bool colok = co < p.Cout && !(u.dbg & 2);
        const bool full4 = co + 3 < p.Cout;
        if (colok && full4 && tile_fast) {
          const float4 b4 = b4c;
          float* yb = p.y + row0 * p.ldy + co;
          const bool sc = p.out_scale != 1.0f;
          if (p.res && p.accumulate) epi_chunk_fast<true, true, true>(stg, lane, acs, p.out_scale, b4, yb, ystep, r4, ps, pq);
          else if (p.res && sc) epi_chunk_fast<true, true, false>(stg, lane, acs, p.out_scale, b4, yb, ystep, r4, ps, pq);
          else if (p.res) epi_chunk_fast<true, false, false>(stg, lane, acs, 1.f, b4, yb, ystep, r4, ps, pq);
          else if (p.accumulate) epi_chunk_fast<false, true, true>(stg, lane, acs, p.out_scale, b4, yb, ystep, r4, ps, pq);
          else if (sc) epi_chunk_fast<false, true, false>(stg, lane, acs, p.out_scale, b4, yb, ystep, r4, ps, pq);
          else epi_chunk_fast<false, false, false>(stg, lane, acs, 1.f, b4, yb, ystep, r4, ps, pq);
        } else if (colok && full4) {
          // all global reads of the chunk first (res / y may alias the output, so the compiler cannot hoist them itself)
          float4 y4[8];
          const float4 b4 = b4c;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int t = tile.t0 + q * 32 + 4 * i + (lane >> 3);
            r4[i] = make_float4(0.f, 0.f, 0.f, 0.f); y4[i] = r4[i];
            if (t < seg.out_len) {
              const size_t orow = (size_t)(seg.out_off + t);
              if (p.res) r4[i] = *reinterpret_cast<const float4*>(p.res + orow * p.ldr + co);
              if (p.accumulate) y4[i] = *reinterpret_cast<const float4*>(p.y + orow * p.ldy + co);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + (lane >> 3);
            const int t = tile.t0 + q * 32 + rr;
            if (t < seg.out_len) {
              const float4 a4 = *reinterpret_cast<const float4*>(stg + rr * STG_LD + cc);
              float4 o;
              o.x = fmaf(a4.x, acs, b4.x) + r4[i].x; o.y = fmaf(a4.y, acs, b4.y) + r4[i].y;
              o.z = fmaf(a4.z, acs, b4.z) + r4[i].z; o.w = fmaf(a4.w, acs, b4.w) + r4[i].w;
              if (p.act == ACT_GELU_NEW) { o.x = gelu_new(o.x); o.y = gelu_new(o.y); o.z = gelu_new(o.z); o.w = gelu_new(o.w); }
              o.x = o.x * p.out_scale + y4[i].x; o.y = o.y * p.out_scale + y4[i].y;
              o.z = o.z * p.out_scale + y4[i].z; o.w = o.w * p.out_scale + y4[i].w;
              *reinterpret_cast<float4*>(p.y + (size_t)(seg.out_off + t) * p.ldy + co) = o;
              ps[0] += o.x; ps[1] += o.y; ps[2] += o.z; ps[3] += o.w;
              pq[0] += o.x * o.x; pq[1] += o.y * o.y; pq[2] += o.z * o.z; pq[3] += o.w * o.w;
            }
          }
        } else if (colok) {
          // ragged last channel group (Cout % 4 != 0): compact scalar loop, deliberately not unrolled (instruction cache)
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + (lane >> 3);
            const int t = tile.t0 + q * 32 + rr;
            if (t < seg.out_len) {
              const size_t orow = (size_t)(seg.out_off + t);
              float* dst = p.y + orow * p.ldy + co;
#pragma unroll 1
              for (int e = 0; e < 4; ++e) {
                if (co + e < p.Cout) {
                  float xv = stg[rr * STG_LD + cc + e] * acs;
                  if (p.bias) xv += p.bias[co + e];
                  if (p.res) xv += p.res[orow * p.ldr + co + e];
                  if (p.act == ACT_GELU_NEW) xv = gelu_new(xv);
                  xv *= p.out_scale;
                  if (p.accumulate) xv += dst[e];
                  dst[e] = xv;
                  if (e == 0) { ps[0] += xv; pq[0] += xv * xv; }
                  else if (e == 1) { ps[1] += xv; pq[1] += xv * xv; }
                  else if (e == 2) { ps[2] += xv; pq[2] += xv * xv; }
                  else { ps[3] += xv; pq[3] += xv * xv; }
                }
              }
            }
          }
        }
        __syncwarp();
        if (p.res && tile_fast && chunk + 1 < nchunk && co + 32 + 3 < p.Cout) {
          const float* rb = p.res + row0 * p.ldr + co + 32;
#pragma unroll
          for (int i = 0; i < 8; ++i) r4[i] = *reinterpret_cast<const float4*>(rb + i * rstep);
        }
        long long tq3 = 0;
        if (tprof) { tq3 = clock64(); tp_io += tq3 - tq2; }
        if (p.ostats) {
          // Column sums over the warp's 32 rows: lanes l, l^8, l^16, l^24 hold partials of the same four columns.  Recursive halving
          // instead of a butterfly: in each round a lane keeps half of its columns and hands the other half to its partner, so 6
          // shuffles do what 16 did and every lane ends with the total of ONE column (cc + 2*bit3 + bit4) - one owner per (row
          // quadrant, channel), two FP64 adds per chunk with all lanes busy.  Same order of additions as the butterfly
          // ((l + l^8) + (l^16 + l^24)): the sums are bit-identical.  The epilogue warps are issue- and MIO-latency bound (profile:
          // ~900 cycles per chunk sat here), so the running sums' old values were requested before the shuffles.
          const bool hi8 = (lane & 8) != 0, hi16 = (lane & 16) != 0;
          const int sidx = q * 256 + c0 + cc + (hi8 ? 2 : 0) + (hi16 ? 1 : 0);
          const double old_s = s_sum[sidx], old_q = s_sq[sidx];
          float ks0 = hi8 ? ps[2] : ps[0], ks1 = hi8 ? ps[3] : ps[1], kq0 = hi8 ? pq[2] : pq[0], kq1 = hi8 ? pq[3] : pq[1];
          ks0 += __shfl_xor_sync(0xffffffffu, hi8 ? ps[0] : ps[2], 8); ks1 += __shfl_xor_sync(0xffffffffu, hi8 ? ps[1] : ps[3], 8);
          kq0 += __shfl_xor_sync(0xffffffffu, hi8 ? pq[0] : pq[2], 8); kq1 += __shfl_xor_sync(0xffffffffu, hi8 ? pq[1] : pq[3], 8);
          float vs = hi16 ? ks1 : ks0, vq = hi16 ? kq1 : kq0;
          vs += __shfl_xor_sync(0xffffffffu, hi16 ? ks0 : ks1, 16);
          vq += __shfl_xor_sync(0xffffffffu, hi16 ? kq0 : kq1, 16);
          s_sum[sidx] = old_s + (double)vs;
          s_sq[sidx] = old_q + (double)vq;
        }
        if (tprof) tp_stats += clock64() - tq3;
      }
      long long tf0 = 0;
      if (tprof) tf0 = clock64();
      if (p.ostats) {
        // flush the per-CTA statistics when the (utterance, Cout tile) changes or the CTA is done
        bool flush = item + 1 >= item1;
        if (!flush) {
          const int nitem = item + 1;
          flush = (nitem / p.ntiles != nt) || (p.tiles[nitem % p.ntiles].seg != tile.seg);
        }
        if (flush) {
          asm volatile("bar.sync 2, %0;" ::"n"(NEPI) : "memory");
          for (int c = et; c < pl.BN; c += NEPI) {
            if (n0 + c < p.Cout) {
              const double ts = (s_sum[c] + s_sum[256 + c]) + (s_sum[512 + c] + s_sum[768 + c]);
              const double tq = (s_sq[c] + s_sq[256 + c]) + (s_sq[512 + c] + s_sq[768 + c]);
              double* st = p.ostats + ((size_t)tile.seg * p.ld_ostats + n0 + c) * 2;
              atomicAdd(st, ts);
              atomicAdd(st + 1, tq);
            }
            s_sum[c] = 0.0; s_sum[256 + c] = 0.0; s_sum[512 + c] = 0.0; s_sum[768 + c] = 0.0;
            s_sq[c] = 0.0; s_sq[256 + c] = 0.0; s_sq[512 + c] = 0.0; s_sq[768 + c] = 0.0;
          }
          asm volatile("bar.sync 2, %0;" ::"n"(NEPI) : "memory");
        }
      }
      if (tprof) tp_post += clock64() - tf0;
    }
    tc_fence_before();
    if (u.prof && et == 0) {
      long long* pp = u.prof + blockIdx.x * 16;
      pp[6] = clock64() - t_start; pp[7] = tw_full; pp[8] = item1 - item0;
      pp[9] = tp_ld; pp[10] = tp_stage; pp[11] = tp_io; pp[12] = tp_stats; pp[13] = tp_pre; pp[14] = tp_top; pp[15] = tp_post;
    }
  }
  __syncthreads();
  if (warp == NPW + 1) tmem_dealloc(tmem_base, (uint32_t)pl.tmem_cols);
}

// Operand images of a layer's input, one per (row tile, K chunk), laid out exactly like the kernel's shared-memory slab
// ([hi | lo][8-channel plane][SR rows][8 x 16 bit], fused prologue applied): a layer with several Cout tiles then bulk-copies
// them instead of converting the same rows again for every Cout tile (the text-side GEMMs and the 1024-channel decoder convs
// were producer bound: up to 18 Cout tiles re-converted each row tile).  Same arithmetic as the in-kernel producers.
template <int PRO, int WFMT>
__global__ void __launch_bounds__(256) pack_act_kernel(const ConvArgs p, int KS, int nsc, int SR, uint32_t slab_bytes, uint8_t* __restrict__ out) {
  __shared__ float s_par[5][128];
  pdl_trigger();
  pdl_wait();                                      // the activations (and the segment tables of the step) come from earlier work
  const Tile tile = p.tiles[blockIdx.x];
  const Seg seg = p.segs[tile.seg];
  const int sc = blockIdx.y;
  const int planes = KS >> 3;
  uint8_t* blk = out + ((size_t)blockIdx.x * nsc + sc) * slab_bytes;
  const size_t lo_off = (size_t)planes * SR * 16;
  if (PRO >= PRO_ADAIN_LRELU) {
    for (int cl = threadIdx.x; cl < KS; cl += blockDim.x) {
      const int c = sc * KS + cl;
      float mean = 0.f, a = 0.f, b = 0.f, al = 1.f;
      if (c < p.Cin) {
        const double* st = p.pstats + ((size_t)tile.seg * p.ld_pstats + c) * 2;
        adain_coef(st, seg.in_len, p.gb[(size_t)tile.seg * p.ld_gb + c], p.gb[(size_t)tile.seg * p.ld_gb + p.beta_off + c], mean, a, b);
        if (PRO == PRO_ADAIN_SNAKE) al = p.alpha[c];
      }
      const float Bf = fmaf(-mean, a, b);
      const float kk = 0.5f / al;
      s_par[0][cl] = a; s_par[1][cl] = (PRO == PRO_ADAIN_SNAKE) ? Bf + kk : Bf;
      s_par[2][cl] = 2.f * al * a; s_par[3][cl] = 2.f * al * Bf; s_par[4][cl] = kk;
    }
    __syncthreads();
  }
  // one item (8 channels of one row) per thread: blockIdx.z covers the rows.  With a loop of 4-8 items per thread every iteration
  // was a global-load round trip of its own (load inside a branch, dependent conversion and store behind it): ~10 us per launch
  // for a few MB, 73 launches per step
  for (int idx = blockIdx.z * blockDim.x + threadIdx.x; idx < planes * SR; idx += gridDim.z * blockDim.x) {
    const int r = idx % SR, pln = idx / SR;
    const int c0 = sc * KS + pln * 8;
    const int traw = tile.t0 + r - p.pad;
    const int tin = traw >> p.in_shr;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (traw >= 0 && tin < seg.in_len && c0 < p.Cin) {
      const float* src = p.x + (size_t)(seg.in_off + tin) * p.ldx + c0;
      if (c0 + 8 <= p.Cin && (p.ldx & 3) == 0) {
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(src)), q1 = __ldg(reinterpret_cast<const float4*>(src + 4));
        v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) if (c0 + e < p.Cin) v[e] = __ldg(src + e);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float x = v[e];
        const int cl = pln * 8 + e;
        if (PRO == PRO_LRELU) x = fmaxf(x, x * p.slope);
        else if (PRO == PRO_ADAIN_LRELU) { const float wv = fmaf(x, s_par[0][cl], s_par[1][cl]); x = fmaxf(wv, wv * p.slope); }
        else if (PRO == PRO_ADAIN_SNAKE) x = fmaf(-s_par[4][cl], snake_cos<WFMT>(fmaf(x, s_par[2][cl], s_par[3][cl])), fmaf(x, s_par[0][cl], s_par[1][cl]));
        v[e] = (c0 + e < p.Cin) ? x : 0.f;
      }
    }
    uint4 o;
    if (WFMT == 2) {
      float hi[8], lo[8];
      float satm = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        satm = fmaxf(satm, fabsf(v[e]));
        const float xs = fminf(fmaxf(v[e] * SPLIT_SA, -65504.f), 65504.f);
        hi[e] = __half2float(__float2half_rn(xs));
        lo[e] = xs - hi[e];
      }
      if (p.sat && satm * SPLIT_SA > 65504.f) atomicAdd(p.sat, 1u);
      uint4 ol;
      o.x = pack_f16_nosat(hi[0], hi[1]); o.y = pack_f16_nosat(hi[2], hi[3]); o.z = pack_f16_nosat(hi[4], hi[5]); o.w = pack_f16_nosat(hi[6], hi[7]);
      ol.x = pack_f16_nosat(lo[0], lo[1]); ol.y = pack_f16_nosat(lo[2], lo[3]); ol.z = pack_f16_nosat(lo[4], lo[5]); ol.w = pack_f16_nosat(lo[6], lo[7]);
      *reinterpret_cast<uint4*>(blk + lo_off + (size_t)idx * 16) = ol;
    } else if (WFMT == 1) {
      if (PRO >= PRO_ADAIN_LRELU) { o.x = pack_f16_nosat(v[0], v[1]); o.y = pack_f16_nosat(v[2], v[3]); o.z = pack_f16_nosat(v[4], v[5]); o.w = pack_f16_nosat(v[6], v[7]); }
      else { o.x = pack_f16(v[0], v[1]); o.y = pack_f16(v[2], v[3]); o.z = pack_f16(v[4], v[5]); o.w = pack_f16(v[6], v[7]); }
    } else { o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]); }
    *reinterpret_cast<uint4*>(blk + (size_t)idx * 16) = o;
  }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }

void pads(int Cin, int Cout, bool split, int& cin_pad, int& cout_pad, int& BN) {
  const int bnmax = split ? 128 : 256;       // split mode streams two weight blocks per stage: halve the Cout tile
  cin_pad = split ? round_up(Cin, 64) : (Cin <= 64 ? 64 : round_up(Cin, 128));
  if (Cout <= bnmax) { cout_pad = round_up(Cout, 32); BN = cout_pad; }
  else { cout_pad = round_up(Cout, bnmax); BN = bnmax; }
}

int g_num_sms[64] = {0};          // per device
bool g_attr_set[64] = {false};    // cudaFuncSetAttribute is per device

bool make_plan(const ConvArgs& a, Plan& pl) {
  int cin_pad, cout_pad, BN;
  pl.split = a.wfmt == 2 ? 1 : 0;
  pads(a.Cin, a.Cout, pl.split != 0, cin_pad, cout_pad, BN);
  if (cin_pad != a.cin_pad || cout_pad != a.cout_pad) return false;
  pl.BN = BN;
  pl.ntile_n = cout_pad / BN;
  pl.KS = pl.split ? 64 : (cin_pad < 128 ? cin_pad : 128);
  pl.nsc = cin_pad / pl.KS;
  pl.kcn = cin_pad / KCH;
  pl.SR = (BM + (a.k - 1) * a.dil) | 1;
  pl.acc_cols = pl.split ? 2 * BN : BN;
  int tc = 32;
  while (tc < 2 * pl.acc_cols) tc <<= 1;           // two accumulator buffers
  pl.tmem_cols = tc;
  uint32_t off = 16;
  pl.off_bar = off; off += 8u * (2 * MAXSTAGE + 8 + 2 * MAXRAW);
  off = (off + 15u) & ~15u;
  pl.off_stats = off; off += 2048 * 8;
  pl.off_par = off; off += a.pro >= PRO_ADAIN_LRELU ? 5u * cin_pad * 4u : 0u;
  off = (off + 15u) & ~15u;
  pl.off_stg = off; off += (NEPI / 32) * STG_BYTES;
  off = (off + 127u) & ~127u;
  pl.off_slab = off;
  pl.slab_bytes = (uint32_t)(pl.KS / 8) * pl.SR * 16u * (pl.split ? 2u : 1u);
  pl.slab_bytes = (pl.slab_bytes + 127u) & ~127u;
  pl.wstage_bytes = (uint32_t)BN * KCH * 2u * (pl.split ? 2u : 1u);
  pl.off_w = pl.off_slab + 2 * pl.slab_bytes;
  // shared memory left after two weight stages goes first to the raw-row ring (>= 2 slots or none), then to more weight stages
  const uint32_t limit = 227u * 1024u;
  pl.nstage = 2;
  pl.nraw = 0; pl.rch = 32; pl.raw_bytes = 0;
  const bool ring_ok = a.in_shr == 0 && (a.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(a.x) & 15) == 0 && getenv("TTSB_UMMA_NORING") == nullptr &&
                       (getenv("TTSB_UMMA_RINGALL") != nullptr || a.ldx == pl.KS);
  if (ring_ok) {
    static const int env_nraw = getenv("TTSB_UMMA_NRAW") ? atoi(getenv("TTSB_UMMA_NRAW")) : 3;
    for (int rch : {2 * NPROD / (pl.KS / 8)}) {        // two rows per producer thread and slot: 16 KB slots
      const uint32_t slot = (uint32_t)rch * pl.KS * 4u;
      const uint32_t used = pl.off_w + 2 * pl.wstage_bytes;
      if (used + 2 * slot <= limit) {
        pl.rch = rch; pl.raw_bytes = slot;
        pl.nraw = (int)std::min<uint32_t>((limit - used) / slot, (uint32_t)std::min(env_nraw, MAXRAW));
        break;
      }
    }
  }
  pl.off_raw = pl.off_w;                       // ring first, weight stages after it
  pl.off_w = pl.off_raw + pl.nraw * pl.raw_bytes;
  while (pl.nstage < MAXSTAGE && pl.off_w + (pl.nstage + 1) * pl.wstage_bytes <= limit) ++pl.nstage;
  pl.smem_bytes = pl.off_w + pl.nstage * pl.wstage_bytes;
  pl.items = a.ntiles * pl.ntile_n;
  return pl.smem_bytes <= 227u * 1024u;
}

}  // namespace

// 16-bit (bf16 or fp16) image of the weights: [Cout tile][tap][64-channel chunk][8-channel plane][BN rows][8 channels]
void pack_umma_weights(const float* w, int k, int Cin, int Cout, int ldw, int fmt, std::vector<uint16_t>& out, int& cin_pad, int& cout_pad,
                       float* split_scale) {
  int BN;
  const bool split = fmt == 2;
  pads(Cin, Cout, split, cin_pad, cout_pad, BN);
  // split mode: the largest power of two that keeps max|w| * scale <= 2^14 (fp16 overflows at 65504; the headroom is for the
  // rounding of the high half), so that the low halves sit as high in fp16's range as the layer allows.  A fixed 256 turned
  // |w| >= 256 into inf and then NaN.
  float SPLIT_SW = 256.f;
  if (split) {
    float mx = 0.f;
    for (int j = 0; j < k; ++j)
      for (int ci = 0; ci < Cin; ++ci)
        for (int co = 0; co < Cout; ++co) mx = std::max(mx, fabsf(w[((size_t)j * Cin + ci) * ldw + co]));
    int ex = 12;
    while (ex > -14 && !(mx * ldexpf(1.f, ex) <= 16384.f)) --ex;      // also walks down for inf / NaN weights (the pack is then garbage either way)
    SPLIT_SW = ldexpf(1.f, ex);
  }
  if (split_scale) *split_scale = SPLIT_SW;
  const int kcn = cin_pad / KCH, ntn = cout_pad / BN;
  const size_t blk = (size_t)BN * KCH * (split ? 2 : 1);          // one stage: [hi block][lo block] in split mode
  out.assign((size_t)ntn * k * kcn * blk, (uint16_t)0);
  for (int nt = 0; nt < ntn; ++nt)
    for (int j = 0; j < k; ++j)
      for (int kc = 0; kc < kcn; ++kc)
        for (int pln = 0; pln < 8; ++pln)
          for (int n = 0; n < BN; ++n) {
            const int co = nt * BN + n;
            if (co >= Cout) continue;
            const size_t base = (((size_t)nt * k + j) * kcn + kc) * blk + ((size_t)pln * BN + n) * 8;
            for (int e = 0; e < 8; ++e) {
              const int ci = kc * KCH + pln * 8 + e;
              if (ci < Cin) {
                const float v = w[((size_t)j * Cin + ci) * ldw + co];
                if (split) {
                  const float vs = v * SPLIT_SW;
                  const __half hv = __float2half(vs);
                  const __half lv = __float2half(vs - __half2float(hv));
                  out[base + e] = *reinterpret_cast<const uint16_t*>(&hv);
                  out[base + (size_t)BN * KCH + e] = *reinterpret_cast<const uint16_t*>(&lv);
                } else if (fmt == 1) { const __half hv = __float2half(v); out[base + e] = *reinterpret_cast<const uint16_t*>(&hv); }
                else { const __nv_bfloat16 bv = __float2bfloat16(v); out[base + e] = *reinterpret_cast<const uint16_t*>(&bv); }
              }
            }
          }
}

bool conv_umma_supported(const ConvArgs& a) {
  if (!a.wb || a.stride != 1 || (a.act != ACT_NONE && a.act != ACT_GELU_NEW)) return false;
  if (a.in_shr != 0 && (a.k != 1 || a.pad != 0)) return false;
  if ((a.ldx & 3) || (reinterpret_cast<uintptr_t>(a.x) & 15)) return false;
  if ((a.ldy & 3) || (reinterpret_cast<uintptr_t>(a.y) & 15)) return false;
  if (a.res && ((a.ldr & 3) || (reinterpret_cast<uintptr_t>(a.res) & 15))) return false;
  if (a.bias && (reinterpret_cast<uintptr_t>(a.bias) & 15)) return false;
  if (a.tile_rows != BM) return false;
  Plan pl;
  return make_plan(a, pl);
}

cudaError_t launch_conv_umma(const ConvArgs& a, cudaStream_t s) {
  if (a.ntiles <= 0) return cudaSuccess;
  UArgs u;
  u.a = a;
  if (!make_plan(a, u.p)) return cudaErrorInvalidValue;
  if (a.wfmt == 2) {
    if (!(a.split_wscale > 0.f)) return cudaErrorInvalidValue;
    u.a.acc_scale = 1.0f / (SPLIT_SA * a.split_wscale);
  }
  using KernelFn = void (*)(const UArgs);
  static const KernelFn table[4][3] = {
      {conv_umma_kernel<PRO_NONE, 0>, conv_umma_kernel<PRO_NONE, 1>, conv_umma_kernel<PRO_NONE, 2>},
      {conv_umma_kernel<PRO_LRELU, 0>, conv_umma_kernel<PRO_LRELU, 1>, conv_umma_kernel<PRO_LRELU, 2>},
      {conv_umma_kernel<PRO_ADAIN_LRELU, 0>, conv_umma_kernel<PRO_ADAIN_LRELU, 1>, conv_umma_kernel<PRO_ADAIN_LRELU, 2>},
      {conv_umma_kernel<PRO_ADAIN_SNAKE, 0>, conv_umma_kernel<PRO_ADAIN_SNAKE, 1>, conv_umma_kernel<PRO_ADAIN_SNAKE, 2>}};
  int dev = 0;
  {
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  }
  if (!g_attr_set[dev]) {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 3; ++j) {
        cudaError_t e = cudaFuncSetAttribute(table[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
      }
    cudaError_t e = cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    g_attr_set[dev] = true;
  }
  if (a.pro < 0 || a.pro > 3 || a.wfmt < 0 || a.wfmt > 2) return cudaErrorInvalidValue;
  static int dbg = -1;
  if (dbg < 0) { const char* ev = getenv("TTSB_UMMA_DBG"); dbg = ev ? atoi(ev) : 0; }
  u.dbg = a.pro == PRO_ADAIN_SNAKE ? dbg : 0;
  static const int l2pf = getenv("TTSB_L2PF") ? atoi(getenv("TTSB_L2PF")) : 0;
  u.l2pf = l2pf & 5;
  static int prof_on = -1;
  static long long* prof_buf = nullptr;
  if (prof_on < 0) {
    prof_on = getenv("TTSB_UMMA_PROF") ? 1 : 0;
    if (prof_on && cudaMalloc(&prof_buf, sizeof(long long) * 16 * 256) != cudaSuccess) prof_on = 0;
  }
  u.prof = prof_on ? prof_buf : nullptr;
  u.xpack = nullptr;
  static const bool nopack = getenv("TTSB_UMMA_NOPACK") != nullptr;
  // worth a pass of its own from three Cout tiles on (measured: at two the pack kernel costs what the producers save)
  if (a.xpack && a.stride == 1 && u.p.ntile_n > 2 && !nopack && (size_t)a.ntiles * u.p.nsc * u.p.slab_bytes <= a.xpack_bytes &&
      (reinterpret_cast<uintptr_t>(a.x) & 15) == 0) {
    using PackFn = void (*)(const ConvArgs, int, int, int, uint32_t, uint8_t*);
    static const PackFn ptable[4][3] = {
        {pack_act_kernel<0, 0>, pack_act_kernel<0, 1>, pack_act_kernel<0, 2>}, {pack_act_kernel<1, 0>, pack_act_kernel<1, 1>, pack_act_kernel<1, 2>},
        {pack_act_kernel<2, 0>, pack_act_kernel<2, 1>, pack_act_kernel<2, 2>}, {pack_act_kernel<3, 0>, pack_act_kernel<3, 1>, pack_act_kernel<3, 2>}};
    const int pack_items = (u.p.KS >> 3) * u.p.SR;
    cudaError_t e = launch_pdl(ptable[a.pro][a.wfmt], dim3(a.ntiles, u.p.nsc, (pack_items + 255) / 256), dim3(256), 0, s, u.a, u.p.KS, u.p.nsc, u.p.SR,
                               u.p.slab_bytes, a.xpack);
    if (e != cudaSuccess) return e;
    e = launched();
    if (e != cudaSuccess) return e;
    u.xpack = a.xpack;
    u.p.nraw = 0;                     // no row ring in this mode
  }
  const KernelFn kernel = table[a.pro][a.wfmt];
  const int ncta = u.p.items < g_num_sms[dev] ? u.p.items : g_num_sms[dev];      // persistent: one CTA per SM
  u.p.per_cta = (u.p.items + ncta - 1) / ncta;
  const int grid = (u.p.items + u.p.per_cta - 1) / u.p.per_cta;
  static cudaEvent_t pe0 = nullptr, pe1 = nullptr;
  if (prof_on) {
    if (!pe0) { cudaEventCreate(&pe0); cudaEventCreate(&pe1); }
    cudaEventRecord(pe0, s);
  }
  {
    cudaError_t e = launch_pdl(kernel, dim3(grid), dim3(NTHREADS), u.p.smem_bytes, s, u);
    if (e != cudaSuccess) return e;
  }
  if (prof_on) cudaEventRecord(pe1, s);
  if (prof_on) {      // diagnostic mode: synchronous, prints the mean cycle split of the roles for every launch
    static long long host[16 * 256];
    if (cudaStreamSynchronize(s) == cudaSuccess && cudaMemcpy(host, prof_buf, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost) == cudaSuccess) {
      double m[16] = {0};
      long long emax = 0, emin = 1LL << 62;
      for (int c = 0; c < grid; ++c) {
        for (int i = 0; i < 16; ++i) m[i] += (double)host[c * 16 + i] / grid;
        emax = host[c * 16 + 6] > emax ? host[c * 16 + 6] : emax;
        emin = host[c * 16 + 6] < emin ? host[c * 16 + 6] : emin;
      }
      float us = 0.f;
      cudaEventElapsedTime(&us, pe0, pe1);
      fprintf(stderr, "umma pro=%d fmt=%d Cin=%d Cout=%d k=%d dil=%d grid=%d items/cta=%.1f | producer %.0f (wait %.0f) | mma %.0f (wait slab %.0f acc %.0f w %.0f) | epilogue %.0f (wait %.0f) cycles | epilogue min %lld max %lld | %.1f us\n",
              a.pro, a.wfmt, a.Cin, a.Cout, a.k, a.dil, grid, m[8], m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], emin, emax, us * 1000.f);
      fprintf(stderr, "   epilogue phases per item (cycles): tmem ld %.0f | staging %.0f | global io + math %.0f | statistics %.0f | item prelude %.0f | chunk top %.0f | flush %.0f | res=%d acc=%d stats=%d\n", m[9] / m[8],
              m[10] / m[8], m[11] / m[8], m[12] / m[8], m[13] / m[8], m[14] / m[8], m[15] / m[8], a.res != nullptr, a.accumulate, a.ostats != nullptr);
    }
  }
  return launched();
}

}  // namespace ttsb
