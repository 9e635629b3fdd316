// tcgen05 implicit-GEMM Conv1d for the AdaINResBlock1 layers of the generator (48 of a step's 144 conv launches, 53 % of the step's
// kernel time: stage 1 with 128 channels, stage 0 with 256), with the WEIGHTS as the A operand:
//
//   D[Cout (M = 128 TMEM lanes), time (N = 256 TMEM columns)] += W[tap][Cout, Cin chunk] * X[time + tap*dil - pad, Cin chunk]^T
//
// Why the roles are swapped relative to conv_umma.cu (written for 128 channels; the 256-channel case is at the end of this comment).
// There the time axis is M (128 rows) and Cout is N; for Cout = 128 an
// M128 x N128 x K16 MMA lasts 64 tensor cycles and reads 4 KB of A and 4 KB of B from shared memory: 128 B/clk, the SM's whole
// shared-memory bandwidth, before the producers, the weight stream and the epilogue staging get any (ncu, round 1:
// sm__pipe_tensor_cycles_active 13-41 % on these layers).  With the weights as A and 256 time rows as N the MMA lasts 128
// cycles for 4 KB + 8 KB = 96 B/clk, the 360-KB weight stream of a k11 layer is re-read once per 256 rows instead of once per
// 128, the halo rows of a slab are 16 % of it instead of 28 %, and the accumulator comes out channel-per-lane: a thread of
// the epilogue owns ONE output channel, so
//   * a warp's store of one time step is 32 consecutive floats (one 128-B line) straight from registers - no shared-memory
//     staging tile, no bank conflicts;
//   * bias is one register, and the per-(utterance, channel) sum / sum of squares for the next AdaIN are per-thread running
//     sums - no shuffles, no shared-memory statistics, one pair of double atomics per thread when the utterance changes.
// The c1 -> c2 intermediate of an AdaINResBlock1 (SURVEY.md A.12 `xt`) is only ever consumed as a 16-bit operand: this kernel
// can write it as fp16 (y_f16) and read it back as fp16 (x_f16), halving its HBM traffic and the raw-row ring's.
//
// Same machinery as conv_umma.cu otherwise: persistent CTAs (one per SM, contiguous tile ranges), warp-specialised roles
// (8 producer warps, weight loader, MMA issuer, 8 epilogue warps, row loader = 608 threads), two operand slabs, a raw-row ring
// fed by the bulk-copy engine, an mbarrier ring of weight stages, two TMEM accumulators (2 x 256 columns = all of TMEM).
// Operand layouts: no-swizzle K-major [8-channel plane][row][16 B] for both; a tap is a row offset of the B descriptor.
// Fused prologue: InstanceNorm normalise + AdaIN affine + Snake (PRO_ADAIN_SNAKE), as conv_umma.cu.
//
// 256-channel layers (stage 0 of the generator; NH = 2).  The same roles work in UNITS of 128 channels: a slab holds one 128-channel
// half h of the tile's input rows, an accumulator one 128-channel half mt of its outputs.  Per tile the MMA warp runs
//     for h: for mt: for tap: D[mt] += W[tap][mt half, h half] X[h half]^T
// so a slab is released after half of the tile's MMAs (the producers stay one unit ahead, as with 128 channels) and the two
// accumulators - all 512 TMEM columns - are staggered: D[0] completes one (unit, mt) pass before D[1], the epilogue drains D[0]
// under the MMAs of D[1] and D[1] under the next tile's first pass.  Units and accumulations are numbered uc = n NH + h and
// g = n NH + mt; their barrier slot is the number's low bit and the phase the next bit - for NH = 1 exactly the two-slot scheme above.
// The weights come in pack order [h][mt][tap][64-channel chunk] (pack_umma_wa_weights), 16 KB per stage as before; the raw-row ring
// takes the unit's 128 channels of each row (one bulk copy per row, the lanes of the row loader side by side).
// When the tiles are fewer than half the SMs (single requests) a work item is (tile, output half mt): two CTAs build the same slabs and
// each runs its half's MMAs - the same sequence per accumulator, so the results are the same bits as in the one-CTA-per-tile mode.
#include "common.cuh"
#include "umma_ptx.cuh"

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

namespace ttsb {

using namespace umma;

namespace {

constexpr int TN = CONV_WA_TILE_ROWS;   // time rows per tile = N of the MMA
constexpr int CH = 128;                 // Cin = Cout
constexpr int KCH = 64;                 // channels per weight stage
constexpr int PLANES = CH / 8;          // 8-channel planes of a slab
constexpr int NPROD = 256;              // 8 producer warps
constexpr int RPP = NPROD / PLANES;     // slab rows the producers cover per pass (16)
constexpr int NEPI = 256;               // 8 epilogue warps: two per TMEM lane quadrant, 128 columns each
constexpr int NTHREADS = NPROD + 64 + NEPI + 32;   // producers (0-7), weight loader (8), MMA issuer (9), epilogue (10-17), row loader (18)
constexpr int MAXRAW = 8;
constexpr int MAXSTAGE = 6;
constexpr uint32_t WSTAGE_BYTES = CH * KCH * 2;    // one (tap, 64-channel) weight stage: 16 KB

struct WPlan {
  int SR;                               // slab rows: TN + (k-1) dil, made odd (conflict-free 16-B stores across planes)
  int nstage, nraw, rps, rch, nch;      // weight stages; ring slots, rows per producer thread per slot, rows per slot, slots per slab
  int items, per_cta;
  int nh;                               // 128-channel halves: 1 (128-channel layers) or 2 (256)
  int split;                            // nh = 2, few tiles: a work item is (tile, output half) - twice the CTAs, half the MMAs each
  uint32_t off_bar, off_par, off_slab, slab_bytes, off_raw, raw_bytes, off_w, smem_bytes, row_bytes;
};

struct WArgs {
  ConvArgs a;
  WPlan p;
  long long* prof;
};

// The epilogue of one tile for one warp: 128 time steps of this thread's output channel.
// The residual (and old-output) rows come in through a ROLLING WINDOW of W registers: row i + W is requested the moment row i has
// been consumed, so every thread has W loads in flight all the time (a warp: W x 128 B).  Measured on B200: with 8 loads per batch
// and nothing overlapped the epilogue ran at the latency of 16 memory round trips per tile (30 k cycles against 11 k of tensor
// time); with units of 16 rows requested one unit ahead, at 8 x ~2.2 k cycles (the load latency under this kernel's own traffic).
// FULL: the half tile lies inside the utterance - straight-line code; otherwise rows past the utterance's end are predicated.
//   y = (acc + bias [+ res]) [* os] [+ y_old]; dsum / dsq += this thread's sum / sum of squares of what it stored
// RES: 0 = no residual, 1 = fp32 residual rows, 2 = fp16 residual rows (the block's running value between its iterations)
template <int RES, bool SCALE, bool ACC, bool YH, bool FULL, int LD>
__device__ __forceinline__ void epi_tile(const uint32_t tacc, const int nv, const float bias, const float os, float* __restrict__ yb, __half* __restrict__ yh,
                                         const void* __restrict__ rbv, const uint32_t bar_empty, const int lane, double& dsum, double& dsq) {
  const float* __restrict__ rb = reinterpret_cast<const float*>(rbv);
  const __half* __restrict__ rh = reinterpret_cast<const __half*>(rbv);
  constexpr size_t ldy = LD, ldr = LD;             // rows of y and res are whole channel rows (conv_umma_wa_supported): the row offsets
                                                   // below are immediates of the load / store instructions, not address registers
  constexpr int W = ACC ? 16 : 32;                 // window depth per stream (two streams when the old output is read too)
  float rw[W], yw[ACC ? W : 1];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    if (RES == 1) rw[i] = (FULL || i < nv) ? rb[(size_t)i * ldr] : 0.f;
    if (RES == 2) rw[i] = (FULL || i < nv) ? __half2float(rh[(size_t)i * ldr]) : 0.f;
    if (ACC) yw[i] = (FULL || i < nv) ? yb[(size_t)i * ldy] : 0.f;
  }
  uint32_t v[16];
  float ps = 0.f, pq = 0.f;
#pragma unroll
  for (int i = 0; i < 128; ++i) {
    if (i % 16 == 0) {
      tmem_ld16(tacc + (uint32_t)i, v);
      if (i == 112) {                              // last read of this accumulator: hand it back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty);
      }
    }
    const bool ok = FULL || i < nv;
    float o = __uint_as_float(v[i % 16]) + bias;
    if (RES) o += rw[i % W];
    if (SCALE && ACC) o = fmaf(o, os, yw[i % W]);
    else if (SCALE) o *= os;
    else if (ACC) o += yw[i % W];
    if (i + W < 128) {                             // refill the slots just consumed: row i + W
      if (RES == 1) rw[i % W] = (FULL || i + W < nv) ? rb[(size_t)(i + W) * ldr] : 0.f;
      if (RES == 2) rw[i % W] = (FULL || i + W < nv) ? __half2float(rh[(size_t)(i + W) * ldr]) : 0.f;
      if (ACC) yw[i % W] = (FULL || i + W < nv) ? yb[(size_t)(i + W) * ldy] : 0.f;
    }
    if (YH) {
      const __half h = __float2half_rn(fminf(fmaxf(o, -65504.f), 65504.f));
      if (ok) yh[(size_t)i * ldy] = h;
      o = __half2float(h);                          // the statistics describe what the consumer will read
    } else {
      if (ok) yb[(size_t)i * ldy] = o;
    }
    if (!FULL) o = ok ? o : 0.f;
    ps += o;
    pq = fmaf(o, o, pq);
    if (i % 16 == 15) { dsum += (double)ps; dsq += (double)pq; ps = 0.f; pq = 0.f; }
  }
}

template <int RES, bool SCALE, bool ACC, bool YH, int LD>
__device__ __forceinline__ void epi_tile_any(const uint32_t tacc, const int nv, const float bias, const float os, float* yb, __half* yh, const void* rb,
                                             const uint32_t bar_empty, const int lane, double& dsum, double& dsq) {
  if (nv >= 128) epi_tile<RES, SCALE, ACC, YH, true, LD>(tacc, 128, bias, os, yb, yh, rb, bar_empty, lane, dsum, dsq);
  else epi_tile<RES, SCALE, ACC, YH, false, LD>(tacc, nv, bias, os, yb, yh, rb, bar_empty, lane, dsum, dsq);
}

template <int WFMT, bool XH, bool YH, int NH>
__global__ void __launch_bounds__(NTHREADS, 1) conv_umma_wa_kernel(const WArgs u) {
  constexpr int CT = NH * CH;                      // channels of the layer (Cin = Cout)
  extern __shared__ __align__(128) uint8_t smem[];
  const ConvArgs& p = u.a;
  const WPlan& pl = u.p;
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
  float* s_par = reinterpret_cast<float*>(smem + pl.off_par);       // [5][CT]: A, B (+k), 2 alpha A, 2 alpha B, k

  const int item0 = blockIdx.x * pl.per_cta;
  const int item1 = min(item0 + pl.per_cta, pl.items);
  // split mode (256 channels, fewer tiles than half the SMs): item = 2 tile + output half.  The accumulation of an output half is the
  // same sequence of MMAs either way (for h: for tap), so the results do not depend on the mode - a request gives the same samples
  // alone (split) and inside a large batch (not split).
  const bool split = NH > 1 && pl.split != 0;
  const int nmt = split ? 1 : NH;                    // output halves per work item
  auto tile_of = [&](int item) { return split ? (item >> 1) : item; };
  auto mt_of = [&](int item) { return split ? (item & 1) : 0; };

  // ---- one-time setup ----
  if (warp == 9) tmem_alloc(smem_u32(s_tmem), 512u);
  if (tid == 0) {
    for (int s = 0; s < MAXSTAGE; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      // one elected lane arrives for its warp (after __syncwarp): 256 lanes arriving on one shared-memory word serialise
      mbar_init(bar_slab_full(b), NPROD / 32); mbar_init(bar_slab_empty(b), 1);
      mbar_init(bar_acc_full(b), 1); mbar_init(bar_acc_empty(b), NEPI / 32);
    }
    for (int b = 0; b < MAXRAW; ++b) { mbar_init(bar_raw_full(b), 1); mbar_init(bar_raw_empty(b), NPROD / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_trigger();                                   // the next kernel on the stream may start its own prologue
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // the weight loader (static data) starts under the previous kernel's tail; every other role waits for that kernel's results
  if (warp != 8) pdl_wait();

  const uint32_t slab0 = smem_u32(smem + pl.off_slab);
  const uint32_t w0 = smem_u32(smem + pl.off_w);
  const int SR = pl.SR;
  constexpr int ESZ = XH ? 2 : 4;
  const uint8_t* xbytes = reinterpret_cast<const uint8_t*>(p.x);

  if (warp < 8) {
    // =================== producers: raw rows -> AdaIN + Snake -> 16-bit operand slab ===================
    const int pln = tid & (PLANES - 1), r0 = tid >> 4;
    const int c0 = pln * 8;
    float pA[8], pB[8], pA2[8], pB2[8], pK[8];
    int cur_seg = -1;
    uint32_t rslot = 0, rphase = 0, n = 0;
    long long t_wait = 0;
    const long long t_start = clock64();
    auto xform_store = [&](uint8_t* slab, int rr, float (&v)[8]) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(-pK[e], __cosf(fmaf(v[e], pA2[e], pB2[e])), fmaf(v[e], pA[e], pB[e]));
      uint4 o;
      // instance-normalised data is bounded (|x^| <= sqrt(n)): no saturation needed for fp16
      if (WFMT == 1) { o.x = pack_f16_nosat(v[0], v[1]); o.y = pack_f16_nosat(v[2], v[3]); o.z = pack_f16_nosat(v[4], v[5]); o.w = pack_f16_nosat(v[6], v[7]); }
      else { o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]); }
      *reinterpret_cast<uint4*>(slab + (size_t)rr * 16) = o;
    };
    auto unpack_h8 = [](const uint4 q, float (&v)[8]) {
      const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
    };
    for (int item = item0; item < item1; ++item, ++n) {
      const Tile tile = p.tiles[tile_of(item)];
      const Seg seg = p.segs[tile.seg];
      const bool new_seg = tile.seg != cur_seg;
      if (new_seg) {
        // per-(utterance, channel) AdaIN coefficients; only producers read them, so a producer-only barrier suffices
        asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
        if (tid < CT) {
          const int c = tid;
          float mean, a, b;
          const double* st = p.pstats + ((size_t)tile.seg * p.ld_pstats + c) * 2;
          adain_coef(st, seg.in_len, p.gb[(size_t)tile.seg * p.ld_gb + c], p.gb[(size_t)tile.seg * p.ld_gb + p.beta_off + c], mean, a, b);
          const float al = p.alpha[c];
          // AdaIN is one FFMA  w = x A + B  (A = (1+gamma) rstd, B = beta - mean A); Snake  w + sin^2(alpha w)/alpha =
          // (w + k) - k cos(2 alpha w),  k = 1/(2 alpha):  wk = x A + (B + k),  t = x (2 alpha A) + 2 alpha B
          const float Bf = fmaf(-mean, a, b);
          const float kk = 0.5f / al;
          s_par[c] = a; s_par[CT + c] = Bf + kk; s_par[2 * CT + c] = 2.f * al * a; s_par[3 * CT + c] = 2.f * al * Bf; s_par[4 * CT + c] = kk;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NPROD) : "memory");
        cur_seg = tile.seg;
      }
      const int row_lo = tile.t0 - p.pad;
#pragma unroll 1
      for (int h = 0; h < NH; ++h) {
        const int cg = h * CH + c0;                 // this thread's eight channels of the layer
        if (NH > 1 || new_seg) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            pA[e] = s_par[cg + e]; pB[e] = s_par[CT + cg + e]; pA2[e] = s_par[2 * CT + cg + e]; pB2[e] = s_par[3 * CT + cg + e]; pK[e] = s_par[4 * CT + cg + e];
          }
        }
        const uint32_t uc = n * NH + h;             // unit number: slab slot = low bit, phase = next bit
        const uint32_t b = uc & 1u;
        mbar_wait_t(bar_slab_empty(b), ((uc >> 1) & 1u) ^ 1u, t_wait);
        uint8_t* slab = smem + pl.off_slab + (size_t)b * pl.slab_bytes + (size_t)pln * SR * 16;
        if (pl.nraw > 0 && row_lo >= 0 && row_lo + SR <= seg.in_len) {
          // interior tile: the row loader streams the unit's raw rows (128 channels each) through the ring with the bulk-copy engine
          for (int ch = 0; ch < pl.nch; ++ch) {
            mbar_wait_t(bar_raw_full(rslot), rphase, t_wait);
            const uint8_t* slot = smem + pl.off_raw + (size_t)rslot * pl.raw_bytes;
            const int ra = ch * pl.rch + r0, rb = ra + RPP;
            const bool hasa = ra < SR, hasb = pl.rps > 1 && rb < SR;
            float va[8], vb[8];
            if (XH) {
              const uint4 qa = *reinterpret_cast<const uint4*>(slot + (size_t)(r0 * CH + c0) * 2);
              const uint4 qb = *reinterpret_cast<const uint4*>(slot + (size_t)((hasb ? r0 + RPP : r0) * CH + c0) * 2);
              unpack_h8(qa, va); unpack_h8(qb, vb);
            } else {
              // groups of four lanes alternate which 16-B half of their 32 B they read first: eight lanes x 16 B at a 32-B
              // stride would hit every bank group twice
              const int hf = (tid >> 2) & 1;
              const float* sa = reinterpret_cast<const float*>(slot) + r0 * CH + c0;
              const float* sb = hasb ? sa + RPP * CH : sa;
              const float4 ax = *reinterpret_cast<const float4*>(sa + 4 * hf), ay = *reinterpret_cast<const float4*>(sa + 4 * (hf ^ 1));
              const float4 bx = *reinterpret_cast<const float4*>(sb + 4 * hf), by = *reinterpret_cast<const float4*>(sb + 4 * (hf ^ 1));
              const float4 a0 = hf ? ay : ax, a1 = hf ? ax : ay, b0 = hf ? by : bx, b1 = hf ? bx : by;
              va[0] = a0.x; va[1] = a0.y; va[2] = a0.z; va[3] = a0.w; va[4] = a1.x; va[5] = a1.y; va[6] = a1.z; va[7] = a1.w;
              vb[0] = b0.x; vb[1] = b0.y; vb[2] = b0.z; vb[3] = b0.w; vb[4] = b1.x; vb[5] = b1.y; vb[6] = b1.z; vb[7] = b1.w;
            }
            if (hasa) xform_store(slab, ra, va);
            if (hasb) xform_store(slab, rb, vb);
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_raw_empty(rslot));      // this warp is done reading the slot
            if (++rslot == (uint32_t)pl.nraw) { rslot = 0; rphase ^= 1u; }
          }
        } else {
          // edge tile of an utterance (or no ring): rows outside the utterance are zero AFTER the activation, as torch pads
          for (int rr = r0; rr < SR; rr += RPP) {
            const int tin = row_lo + rr;
            float v[8];
            if (tin >= 0 && tin < seg.in_len) {
              const uint8_t* src = xbytes + ((size_t)(seg.in_off + tin) * p.ldx + cg) * ESZ;
              if (XH) unpack_h8(__ldg(reinterpret_cast<const uint4*>(src)), v);
              else {
                const float4 q0 = __ldg(reinterpret_cast<const float4*>(src)), q1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
                v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
              }
              xform_store(slab, rr, v);
            } else {
              *reinterpret_cast<uint4*>(slab + (size_t)rr * 16) = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_slab_full(b));
      }
    }
    if (u.prof && tid == 0) { u.prof[blockIdx.x * 16 + 0] = clock64() - t_start; u.prof[blockIdx.x * 16 + 1] = t_wait; }
  } else if (warp == 8) {
    // =================== weight loader (bulk-copy engine) ===================
    if (lane == 0) {
      uint32_t it = 0;
      // stages in pack order [h][mt][tap][64-channel chunk]; an item takes the output halves [mt0, mt0 + nmt) of both input halves
      for (int item = item0; item < item1; ++item) {
        const int mt0 = mt_of(item);
        for (int h = 0; h < NH; ++h)
          for (int mt = mt0; mt < mt0 + nmt; ++mt)
            for (int jk = 0; jk < p.k * (CH / KCH); ++jk, ++it) {
              const uint32_t s = it % NS;
              const size_t blk = (size_t)((h * NH + mt) * p.k * (CH / KCH) + jk);
              mbar_wait(bar_empty(s), ((it / NS) & 1u) ^ 1u);
              mbar_expect_tx(bar_full(s), WSTAGE_BYTES);
              bulk_g2s(w0 + s * WSTAGE_BYTES, p.wb + blk * (WSTAGE_BYTES / 2), WSTAGE_BYTES, bar_full(s));
            }
      }
    }
  } else if (warp == 9) {
    // =================== MMA issuer ===================
    // the whole warp walks the loops with warp-uniform values (descriptors live in uniform registers); one elected lane issues
    uint32_t leader;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(leader));
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t ofmt = WFMT >= 1 ? 0u : 1u;       // operand format: 0 = F16, 1 = BF16
    const uint32_t idesc = (1u << 4) | (ofmt << 7) | (ofmt << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(CH >> 4) << 24);
    // A = weights [8-ch plane][128 Cout rows][16 B]: LBO (between the two K chunks of a K16 step) = 128 * 16, SBO (8-row groups) = 128
    // B = slab    [8-ch plane][SR rows][16 B]:       LBO = SR * 16, SBO = 128; a tap is a row offset of the start address
    const uint64_t a_desc0 = make_desc(0, (uint32_t)CH * 16u, 128u), b_desc0 = make_desc(0, (uint32_t)SR * 16u, 128u);
    const uint32_t a_kstep = 2u * CH;                  // two planes per K16 step, in 16-B units
    const uint32_t b_kstep = 2u * (uint32_t)SR;
    uint32_t it = 0, n = 0;
    long long tw_slab = 0, tw_acc = 0, tw_w = 0;
    const long long t_start = clock64();
    for (int item = item0; item < item1; ++item, ++n) {
      const int mt0 = mt_of(item);
#pragma unroll 1
      for (int h = 0; h < NH; ++h) {
        const uint32_t uc = n * NH + h;                  // unit: one 128-channel half of the tile's input rows
        const uint32_t sl = uc & 1u;
        mbar_wait_t(bar_slab_full(sl), (uc >> 1) & 1u, tw_slab);
        tc_fence_after();
        const uint32_t sbase = slab0 + sl * pl.slab_bytes;
#pragma unroll 1
        for (int mt = mt0; mt < mt0 + nmt; ++mt) {
          const uint32_t g = n * (uint32_t)nmt + (uint32_t)(mt - mt0);      // accumulation: one 128-channel half of the tile's outputs
          const uint32_t slot = g & 1u;
          if (h == 0) {
            mbar_wait_t(bar_acc_empty(slot), ((g >> 1) & 1u) ^ 1u, tw_acc);       // epilogue has drained this accumulator
            tc_fence_after();
          }
          const uint32_t tacc = tmem_u + slot * (uint32_t)TN;
          uint32_t first = h == 0 ? 1u : 0u;
          for (int j = 0; j < p.k; ++j)
            for (int kc = 0; kc < CH / KCH; ++kc, ++it) {
              const uint32_t s = it % NS;
              mbar_wait_t(bar_full(s), (it / NS) & 1u, tw_w);
              tc_fence_after();
              const uint64_t ad0 = a_desc0 + (uint64_t)((w0 + s * WSTAGE_BYTES) >> 4);
              const uint64_t bd0 = b_desc0 + (uint64_t)((sbase >> 4) + (uint32_t)(kc * 8 * SR + j * p.dil));
#pragma unroll
              for (int kk = 0; kk < KCH / 16; ++kk) {
                if (leader) umma_bf16(tacc, ad0 + (uint64_t)(kk * a_kstep), bd0 + (uint64_t)(kk * b_kstep), idesc, first ? 0u : 1u);
                first = 0;
              }
              if (leader) umma_commit(bar_empty(s));            // weight stage free once these MMAs have read it
              __syncwarp();
            }
          if (h == NH - 1) {
            if (leader) umma_commit(bar_acc_full(slot));        // this half of the tile's outputs is complete
            __syncwarp();
          }
        }
        if (leader) umma_commit(bar_slab_empty(sl));            // slab free once every MMA issued so far has completed
        __syncwarp();
      }
    }
    if (u.prof && lane == 0) {
      long long* pp = u.prof + blockIdx.x * 16;
      pp[2] = clock64() - t_start; pp[3] = tw_slab; pp[4] = tw_acc; pp[5] = tw_w;
    }
  } else if (warp == 18) {
    // =================== row loader: raw activation rows -> ring (bulk-copy engine) ===================
    // (An L2 prefetch of the next tile's rows from here was measured slower - it doubles the request traffic of a memory system that
    //  is throughput-bound under this kernel - and was removed; profiles/r02_bench_log.md, row 7.)
    uint32_t rslot = 0, rphase = 0;
    const uint32_t ubytes = (uint32_t)CH * ESZ;              // a unit's share of one row
    for (int item = item0; item < item1; ++item) {
      if (pl.nraw == 0) continue;
      const Tile tile = p.tiles[tile_of(item)];
      const Seg seg = p.segs[tile.seg];
      const int row_lo = tile.t0 - p.pad;
      if (!(row_lo >= 0 && row_lo + SR <= seg.in_len)) continue;      // edge tile: the producers load it themselves
#pragma unroll 1
      for (int h = 0; h < NH; ++h)
        for (int ch = 0; ch < pl.nch; ++ch) {
          if (lane == 0) mbar_wait(bar_raw_empty(rslot), rphase ^ 1u);
          __syncwarp();
          const int rows_n = min(pl.rch, SR - ch * pl.rch);
          const uint8_t* src = xbytes + ((size_t)(seg.in_off + row_lo + ch * pl.rch) * CT + (size_t)h * CH) * ESZ;
          const uint32_t dst = smem_u32(smem + pl.off_raw + (size_t)rslot * pl.raw_bytes);
          if (lane == 0) mbar_expect_tx(bar_raw_full(rslot), (uint32_t)rows_n * ubytes);
          __syncwarp();                                    // the byte count is registered before any copy can complete
          if (NH == 1) {                                   // ldx == CH: the rows are one contiguous block
            if (lane == 0) bulk_g2s(dst, src, (uint32_t)rows_n * ubytes, bar_raw_full(rslot));
          } else {                                         // 128 of a row's 256 channels: one copy per row, the lanes side by side (rch <= 32)
            if (lane < rows_n) bulk_g2s(dst + (uint32_t)lane * ubytes, src + (size_t)lane * CT * ESZ, ubytes, bar_raw_full(rslot));
          }
          if (++rslot == (uint32_t)pl.nraw) { rslot = 0; rphase ^= 1u; }
        }
    }
  } else {
    // =================== epilogue: 8 warps, TMEM lane quadrant q = warp % 4, columns [128 half, 128 half + 128) ===================
    const int q = warp & 3;
    const int half = (warp - 10) >> 2;
    const int cl = q * 32 + lane;                      // this thread's output channel within a 128-channel half
    // per-thread state of the (up to) two output halves in scalars, selected by mt: the loop over mt stays rolled (one copy of the
    // nine epilogue variants) and nothing lives in local memory
    const float bias0 = p.bias ? p.bias[cl] : 0.f, bias1 = (p.bias && NH > 1) ? p.bias[CH + cl] : 0.f;
    double dsum0 = 0.0, dsq0 = 0.0, dsum1 = 0.0, dsq1 = 0.0;
    const bool has_res = p.res != nullptr, acc = p.accumulate != 0, sc = p.out_scale != 1.0f;
    constexpr size_t ldy = CT, ldr = CT;
    uint32_t n = 0;
    long long tw_full = 0;
    const long long t_start = clock64();
    for (int item = item0; item < item1; ++item, ++n) {
      const Tile tile = p.tiles[tile_of(item)];
      const Seg seg = p.segs[tile.seg];
      const int mt0 = mt_of(item);
      const int tbase = tile.t0 + half * 128;
      const int nv = seg.out_len - tbase;              // rows of this half tile inside the utterance (<= 0: nothing to store)
      const size_t row0 = (size_t)(seg.out_off + tbase);
#pragma unroll 1
      for (int mt = mt0; mt < mt0 + nmt; ++mt) {
        const uint32_t g = n * (uint32_t)nmt + (uint32_t)(mt - mt0);        // accumulation number: slot = low bit, phase = next bit
        const uint32_t slot = g & 1u;
        const int c = mt * CH + cl;
        const float bs = mt ? bias1 : bias0;
        double ds = mt ? dsum1 : dsum0, dq = mt ? dsq1 : dsq0;
        mbar_wait_t(bar_acc_full(slot), (g >> 1) & 1u, tw_full);
        tc_fence_after();
        const uint32_t tacc = tmem_base + slot * (uint32_t)TN + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 128);
        float* yb = YH ? nullptr : p.y + row0 * ldy + c;
        __half* yh = YH ? reinterpret_cast<__half*>(p.y) + row0 * ldy + c : nullptr;
        // residual rows: fp32 (the stage's input, first iteration of a block) or fp16 (the block's running value)
        const void* rb = !has_res ? nullptr
                         : p.r_f16 ? static_cast<const void*>(reinterpret_cast<const __half*>(p.res) + row0 * ldr + c)
                                   : static_cast<const void*>(p.res + row0 * ldr + c);
        const uint32_t be = bar_acc_empty(slot);
        if (YH) {                        // c1 (no residual) and the c2 of a block's first two iterations (running value out as fp16)
          if (!has_res) epi_tile_any<0, false, false, true, CT>(tacc, nv, bs, 1.f, yb, yh, rb, be, lane, ds, dq);
          else if (p.r_f16) epi_tile_any<2, false, false, true, CT>(tacc, nv, bs, 1.f, yb, yh, rb, be, lane, ds, dq);
          else epi_tile_any<1, false, false, true, CT>(tacc, nv, bs, 1.f, yb, yh, rb, be, lane, ds, dq);
        } else if (p.r_f16) {            // the c2 of a block's last iteration: fp32 out (scaled, accumulated over the parallel blocks)
          if (acc) epi_tile_any<2, true, true, false, CT>(tacc, nv, bs, p.out_scale, yb, yh, rb, be, lane, ds, dq);
          else if (sc) epi_tile_any<2, true, false, false, CT>(tacc, nv, bs, p.out_scale, yb, yh, rb, be, lane, ds, dq);
          else epi_tile_any<2, false, false, false, CT>(tacc, nv, bs, 1.f, yb, yh, rb, be, lane, ds, dq);
        } else {
          if (acc) epi_tile_any<1, true, true, false, CT>(tacc, nv, bs, p.out_scale, yb, yh, rb, be, lane, ds, dq);
          else if (sc) epi_tile_any<1, true, false, false, CT>(tacc, nv, bs, p.out_scale, yb, yh, rb, be, lane, ds, dq);
          else epi_tile_any<1, false, false, false, CT>(tacc, nv, bs, 1.f, yb, yh, rb, be, lane, ds, dq);
        }
        if (mt) { dsum1 = ds; dsq1 = dq; } else { dsum0 = ds; dsq0 = dq; }
      }
      if (p.ostats) {
        // flush this thread's running sums when the utterance changes or the CTA is done
        const bool flush = item + 1 >= item1 || p.tiles[tile_of(item + 1)].seg != tile.seg;
        if (flush) {
          double* st = p.ostats + ((size_t)tile.seg * p.ld_ostats + cl) * 2;
          atomicAdd(st, dsum0);
          atomicAdd(st + 1, dsq0);
          if (NH > 1) {
            atomicAdd(st + 2 * CH, dsum1);
            atomicAdd(st + 2 * CH + 1, dsq1);
          }
          dsum0 = 0.0; dsq0 = 0.0; dsum1 = 0.0; dsq1 = 0.0;
        }
      }
    }
    tc_fence_before();
    if (u.prof && tid == NPROD + 64) { u.prof[blockIdx.x * 16 + 6] = clock64() - t_start; u.prof[blockIdx.x * 16 + 7] = tw_full; u.prof[blockIdx.x * 16 + 8] = item1 - item0; }
  }
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512u);
}

bool make_wplan(const ConvArgs& a, WPlan& pl) {
  if ((a.Cin != CH && a.Cin != 2 * CH) || a.Cout != a.Cin || a.cin_pad != a.Cin || a.cout_pad != a.Cin) return false;
  pl.nh = a.Cin / CH;
  pl.SR = (TN + (a.k - 1) * a.dil) | 1;
  pl.row_bytes = (uint32_t)CH * (a.x_f16 ? 2u : 4u);      // a unit's share of one row
  uint32_t off = 16;
  pl.off_bar = off; off += 8u * (2 * MAXSTAGE + 8 + 2 * MAXRAW);
  off = (off + 15u) & ~15u;
  pl.off_par = off; off += 5u * (uint32_t)a.Cin * 4u;
  off = (off + 127u) & ~127u;
  pl.off_slab = off;
  pl.slab_bytes = ((uint32_t)PLANES * pl.SR * 16u + 127u) & ~127u;
  off += 2 * pl.slab_bytes;
  const uint32_t limit = 227u * 1024u;
  // three weight stages (3 x 4 MMAs x 128 cycles of cover for one L2 round trip), then the ring, then more weight stages
  pl.nstage = 3;
  // experiment knobs: fewer weight stages (more ring) for the fp32-input layers with few taps, where the ring starves the producers
  // (measured: k3 with fp32 input 236 -> 201 us with 2 stages and a third 16-kB ring slot - a k3 tile has 6 weight stages and
  //  3 k cycles of MMAs in 12 k; k7 unchanged either way)
  static const int env_ns = getenv("TTSB_WA_NSTAGE") ? atoi(getenv("TTSB_WA_NSTAGE")) : 2;
  static const int env_kmax = getenv("TTSB_WA_NSTAGE_KMAX") ? atoi(getenv("TTSB_WA_NSTAGE_KMAX")) : 3;
  static const int env_rps = getenv("TTSB_WA_RPS") ? atoi(getenv("TTSB_WA_RPS")) : 0;
  if (!a.x_f16 && a.k <= env_kmax && env_ns >= 2 && env_ns <= 3) pl.nstage = env_ns;
  const int nstage_min = pl.nstage;
  if (off + pl.nstage * WSTAGE_BYTES > limit) return false;
  const uint32_t left = limit - off - pl.nstage * WSTAGE_BYTES;
  static const bool noring = getenv("TTSB_UMMA_NORING") != nullptr;
  // a slot holds rps rows per producer thread: two rows (one barrier round trip per two rows: 16 KB of fp32, 8 KB of fp16) when
  // at least two such slots fit, else one
  pl.rps = 2;
  pl.raw_bytes = (uint32_t)(2 * RPP) * pl.row_bytes;
  if (left / pl.raw_bytes < 2 || (env_rps == 1 && !a.x_f16)) { pl.rps = 1; pl.raw_bytes = (uint32_t)RPP * pl.row_bytes; }
  pl.rch = pl.rps * RPP;
  pl.nraw = (int)std::min<uint32_t>(left / pl.raw_bytes, (uint32_t)MAXRAW);
  if (pl.nraw < 2 || noring) pl.nraw = 0;
  pl.nch = (pl.SR + pl.rch - 1) / pl.rch;
  pl.off_raw = off; off += pl.nraw * pl.raw_bytes;
  pl.off_w = off;
  while (pl.nstage < MAXSTAGE && pl.off_w + (pl.nstage + 1) * WSTAGE_BYTES <= limit) ++pl.nstage;
  (void)nstage_min;
  pl.smem_bytes = pl.off_w + pl.nstage * WSTAGE_BYTES;
  pl.items = a.ntiles;
  pl.split = 0;
  return pl.smem_bytes <= limit;
}

int g_num_sms_wa[64] = {0};
bool g_attr_set_wa[64] = {false};

}  // namespace

// The weights of a Cin = Cout = 128 NH layer in the order the MMA warp consumes them: [h][mt][tap][64-channel chunk], each stage
// [8-channel plane][128 Cout rows][8 channels] (no-swizzle K-major, the A operand).  w: [k][C][ldw] fp32.  fmt: 0 = bf16, 1 = fp16.
void pack_umma_wa_weights(const float* w, int k, int C, int ldw, int fmt, std::vector<uint16_t>& out) {
  const int nh = C / CH;
  out.assign((size_t)nh * nh * k * (CH / KCH) * CH * KCH, (uint16_t)0);
  size_t base = 0;
  for (int h = 0; h < nh; ++h)
    for (int mt = 0; mt < nh; ++mt)
      for (int j = 0; j < k; ++j)
        for (int kc = 0; kc < CH / KCH; ++kc, base += (size_t)CH * KCH)
          for (int pln = 0; pln < 8; ++pln)
            for (int n = 0; n < CH; ++n)
              for (int e = 0; e < 8; ++e) {
                const int ci = h * CH + kc * KCH + pln * 8 + e, co = mt * CH + n;
                const float v = w[((size_t)j * C + ci) * ldw + co];
                uint16_t bits;
                if (fmt == 1) { const __half hv = __float2half(v); bits = *reinterpret_cast<const uint16_t*>(&hv); }
                else { const __nv_bfloat16 bv = __float2bfloat16(v); bits = *reinterpret_cast<const uint16_t*>(&bv); }
                out[base + ((size_t)pln * CH + n) * 8 + e] = bits;
              }
}

bool conv_umma_wa_supported(const ConvArgs& a) {
  if (!a.wb || a.wfmt < 0 || a.wfmt > 1 || a.pro != PRO_ADAIN_SNAKE || a.stride != 1 || a.in_shr != 0 || a.act != ACT_NONE) return false;
  if (a.tile_rows != TN || a.ldx != a.Cin || a.ldy != a.Cin || (a.res && a.ldr != a.Cin)) return false;
  if (a.acc_scale != 0.f && a.acc_scale != 1.0f) return false;
  if (reinterpret_cast<uintptr_t>(a.x) & 15) return false;
  if (a.y_f16 && (a.accumulate || a.out_scale != 1.0f)) return false;
  if (!a.y_f16 && !a.res) return false;        // the shapes of an AdaINResBlock1: c1 (fp16 out, plain), c2 (+ residual; fp16 or fp32 out)
  if (a.r_f16 && !a.res) return false;
  if (a.res && (reinterpret_cast<uintptr_t>(a.res) & 1)) return false;
  if (!a.pstats || !a.gb || !a.alpha) return false;
  WPlan pl;
  return make_wplan(a, pl);
}

cudaError_t launch_conv_umma_wa(const ConvArgs& a, cudaStream_t s) {
  if (a.ntiles <= 0) return cudaSuccess;
  if (!conv_umma_wa_supported(a)) return cudaErrorInvalidValue;
  WArgs u;
  u.a = a;
  if (!make_wplan(a, u.p)) return cudaErrorInvalidValue;
  using KernelFn = void (*)(const WArgs);
  // [channels / 128 - 1][wfmt][x_f16][y_f16]: fp32 in / fp16 out = c1 of a block's first iteration; fp16 / fp16 = the other c1s and
  // the c2s that write the block's running value; fp16 in / fp32 out = the last c2; fp32 / fp32 does not occur
  static const KernelFn table[2][2][2][2] = {
      {{{nullptr, conv_umma_wa_kernel<0, false, true, 1>}, {conv_umma_wa_kernel<0, true, false, 1>, conv_umma_wa_kernel<0, true, true, 1>}},
       {{nullptr, conv_umma_wa_kernel<1, false, true, 1>}, {conv_umma_wa_kernel<1, true, false, 1>, conv_umma_wa_kernel<1, true, true, 1>}}},
      {{{nullptr, conv_umma_wa_kernel<0, false, true, 2>}, {conv_umma_wa_kernel<0, true, false, 2>, conv_umma_wa_kernel<0, true, true, 2>}},
       {{nullptr, conv_umma_wa_kernel<1, false, true, 2>}, {conv_umma_wa_kernel<1, true, false, 2>, conv_umma_wa_kernel<1, true, true, 2>}}}};
  const KernelFn kernel = table[u.p.nh - 1][a.wfmt][a.x_f16 ? 1 : 0][a.y_f16 ? 1 : 0];
  if (!kernel) return cudaErrorInvalidValue;
  int dev = 0;
  {
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  }
  if (!g_attr_set_wa[dev]) {
    for (int i = 0; i < 16; ++i) {
      const KernelFn f = (&table[0][0][0][0])[i];
      if (f) {
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
      }
    }
    cudaError_t e = cudaDeviceGetAttribute(&g_num_sms_wa[dev], cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    g_attr_set_wa[dev] = true;
  }
  static int prof_on = -1;
  static long long* prof_buf = nullptr;
  if (prof_on < 0) {
    prof_on = getenv("TTSB_UMMA_PROF") ? 1 : 0;
    if (prof_on && cudaMalloc(&prof_buf, sizeof(long long) * 16 * 256) != cudaSuccess) prof_on = 0;
  }
  u.prof = prof_on ? prof_buf : nullptr;
  // few 256-channel tiles (single requests, short batches): one work item per (tile, output half), so that twice as many SMs share
  // the MMAs (a k11 tile is 45 k cycles of them); bit-identical results (see the kernel)
  static const bool no_split = getenv("TTSB_WA_NOSPLIT") != nullptr;
  if (u.p.nh == 2 && !no_split && 2 * u.p.items <= g_num_sms_wa[dev]) { u.p.split = 1; u.p.items *= 2; }
  const int ncta = u.p.items < g_num_sms_wa[dev] ? u.p.items : g_num_sms_wa[dev];      // persistent: one CTA per SM
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
  if (prof_on) {      // diagnostic mode: synchronous, prints the mean cycle split of the roles for every launch
    cudaEventRecord(pe1, s);
    static long long host[16 * 256];
    if (cudaStreamSynchronize(s) == cudaSuccess && cudaMemcpy(host, prof_buf, sizeof(long long) * 16 * grid, cudaMemcpyDeviceToHost) == cudaSuccess) {
      double m[9] = {0};
      for (int cta = 0; cta < grid; ++cta)
        for (int i = 0; i < 9; ++i) m[i] += (double)host[cta * 16 + i] / grid;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, pe0, pe1);
      fprintf(stderr, "umma_wa C=%d fmt=%d xh=%d yh=%d k=%d dil=%d grid=%d items/cta=%.1f SR=%d ring=%dx%uB wstages=%d | producer %.0f (wait %.0f) | mma %.0f (wait slab %.0f acc %.0f w %.0f) | epilogue %.0f (wait %.0f) cycles | %.1f us\n",
              a.Cin, a.wfmt, a.x_f16, a.y_f16, a.k, a.dil, grid, m[8], u.p.SR, u.p.nraw, u.p.raw_bytes, u.p.nstage, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], ms * 1000.f);
    }
  }
  return launched();
}

}  // namespace ttsb
