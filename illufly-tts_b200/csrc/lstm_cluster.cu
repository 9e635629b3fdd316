// Persistent cluster BiLSTM recurrence (hidden 256 per direction), fp32.
//
// One thread-block CLUSTER of 8 CTAs runs one direction of the recurrence for a group of LSTM_UB utterances from the first
// step to the last without leaving the SMs:
//   * W_hh of the direction (1024 x 256 fp32 = 1 MB) lives in REGISTERS for the whole kernel: CTA `rank` owns the four gate
//     rows of hidden units [32 rank, 32 rank + 32) (128 rows); a thread owns a 4-row x 16-k block (64 weights), so every
//     16-B load of h feeds 16 FMAs, and the 16 k-slices of a row group are reduced with a 15-shuffle reduce-scatter;
//   * h_t of every utterance of the group lives in shared memory of every CTA (double buffered); after the cell update each
//     CTA pushes its 32 new hidden values per utterance into all 8 CTAs with st.async (STAS): a remote shared-memory store that
//     completes bytes on the RECEIVER's mbarrier, so a step needs no fence and no cluster barrier - every thread just waits
//     on its own CTA's mbarrier for the bytes of h_t.  (A barrier.cluster.arrive.release per step compiled to MEMBAR.ALL.GPU
//     in all 16 warps, which also waited for the global y store, and the .acquire wait to CCTL.IVALL: ncu, 40 % of the kernel.)
//     Double buffering needs no extra guard: a peer can only push h_{t+2} into the buffer h_t sits in once it holds all of
//     h_{t+1}, which every CTA sent after it had finished reading h_t;
//   * gx = x W_ih^T + b (computed by the GEMM kernel for all steps at once) is prefetched one step ahead.
// Gate order i,f,g,o and c' = f c + i g, h = o tanh(c') as torch.nn.LSTM.  The reverse direction may start
// with `pre` zero-input steps (the un-packed predictor.lstm of forward_batch, kmodel.py:95).
#include "common.cuh"

namespace ttsb {

namespace {

constexpr int NCTA = 8;             // cluster size: 8 x 32 hidden units
constexpr int NT = 512;             // 32 row groups (4 rows) x 16 k-slices (16 k)
static_assert(LSTM_UB == 4, "the reduce-scatter below is written for 4 rows x 4 utterances = 16 values over 16 lanes");

// h is stored permuted so that the 16 k-slices read consecutive 16-B words: k = 16 ks + 4 i + e  ->  i*64 + ks*4 + e
__device__ __forceinline__ int hpos(int k) { return ((k & 15) >> 2) * 64 + (k >> 4) * 4 + (k & 3); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// two independent IEEE fp32 FMAs in one instruction (SASS FFMA2): (a0, a1) = (w0 h0 + a0, w1 h1 + a1), bit-identical to two fmaf
__device__ __forceinline__ void fma2(float& a0, float& a1, float w0, float w1, float h0, float h1) {
  unsigned long long a, w, h;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(w0), "f"(w1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(h) : "f"(h0), "f"(h1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(w), "l"(h));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}
__device__ __forceinline__ uint32_t map_to_peer(uint32_t local_smem_addr, uint32_t peer) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_smem_addr), "r"(peer));
  return remote;
}
// remote store of one float that completes 4 bytes on the receiver's mbarrier (no fence on either side)
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) __trap();
  }
}

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(NT, 1)
lstm_cluster_kernel(const float* __restrict__ gx, const float* __restrict__ whh, const float* __restrict__ bias, float* __restrict__ y, int ldy,
                    const LstmGroup* __restrict__ groups) {
  __shared__ __align__(16) float hbuf[2][LSTM_UB][256];
  __shared__ float gates_s[LSTM_UB][128];
  __shared__ __align__(8) unsigned long long hbar_s[2];         // hbar_s[b]: the bytes of the h that lands in hbuf[b]
  const int rank = (int)cluster_ctarank();
  const LstmGroup grp = groups[blockIdx.y];
  const int dir = blockIdx.z;
  const int tid = threadIdx.x;
  const int rg = tid >> 4, ks = tid & 15;

  // this thread's 4 x 16 block of recurrent weights, registers for the whole kernel
  float w[4][16];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row_l = 4 * rg + r;                                  // local gate row: gate * 32 + hidden unit
    const int grow = (row_l >> 5) * 256 + rank * 32 + (row_l & 31);
    const float4* src = reinterpret_cast<const float4*>(whh + ((size_t)dir * 1024 + grow) * 256 + ks * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = src[i];
      w[r][4 * i] = t.x; w[r][4 * i + 1] = t.y; w[r][4 * i + 2] = t.z; w[r][4 * i + 3] = t.w;
    }
  }
  for (int i = tid; i < 2 * LSTM_UB * 256; i += NT) (&hbuf[0][0][0])[i] = 0.f;

  // cell threads: (utterance u, hidden unit cj) for the first 32 * LSTM_UB threads
  const bool cell = tid < 32 * LSTM_UB;
  const int cu = tid >> 5, cj = tid & 31;
  int c_off = 0, c_len = 0, c_pre = 0, c_steps = 0;
  if (cell) { c_off = grp.off[cu]; c_len = grp.len[cu]; c_pre = dir ? grp.pre[cu] : 0; c_steps = c_len > 0 ? c_len + c_pre : 0; }
  int maxsteps = 0;
  int usteps[LSTM_UB];                                   // steps of every utterance of the group: the same in all 8 CTAs
#pragma unroll
  for (int u = 0; u < LSTM_UB; ++u) {
    const int st = grp.len[u] > 0 ? grp.len[u] + (dir ? grp.pre[u] : 0) : 0;
    usteps[u] = st;
    maxsteps = st > maxsteps ? st : maxsteps;
  }
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&hbar_s[0]);       // hbar_s[b] at bar0 + 8 b
  const uint32_t hb0 = (uint32_t)__cvta_generic_to_shared(&hbuf[0][0][0]);    // hbuf[b] at hb0 + 4096 b
  const int hcol = rank * 32 + cj;                       // my hidden unit (index within the direction)
  const int hidx = hpos(hcol);
  float bg[4] = {0.f, 0.f, 0.f, 0.f};
  if (cell) {
#pragma unroll
    for (int g = 0; g < 4; ++g) bg[g] = bias[dir * 1024 + g * 256 + hcol];
  }
  float c = 0.f;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_arrive();
  cluster_wait();            // once: every CTA's hbuf is zeroed and its mbarriers exist before anybody pushes into it
  if (tid == 0) mbar_arrive(bar0);       // h_0 = 0 already sits in hbuf[0]: its phase completes with nothing to receive

  for (int s = 0; s < maxsteps; ++s) {
    const int cur = s & 1;
    if (tid == 0) {
      // arm the receiver for h_{s+1}: 32 hidden units x 4 bytes from each of the 8 CTAs for every utterance still running.
      // (The phase this re-arms, h_{s-1}, was observed by every thread of this CTA: they all passed __syncthreads of step s-1.)
      int n_act = 0;
#pragma unroll
      for (int u = 0; u < LSTM_UB; ++u) n_act += (s < usteps[u]) ? 1 : 0;
      mbar_arrive_expect_tx(bar0 + 8u * (uint32_t)(cur ^ 1), (uint32_t)n_act * 256u * 4u);
    }
    // prefetch this step's input gates (latency hidden behind the dot products)
    float gxi[4] = {0.f, 0.f, 0.f, 0.f};
    bool active = false, is_pad = false;
    int t = 0;
    if (cell) {
      active = s < c_steps;
      is_pad = s < c_pre;
      t = dir ? (c_len - 1 - (s - c_pre)) : s;
      if (active) {
        if (is_pad) {
#pragma unroll
          for (int g = 0; g < 4; ++g) gxi[g] = bg[g];
        } else {
          const float* gp = gx + (size_t)(c_off + t) * 2048 + dir * 1024 + hcol;
#pragma unroll
          for (int g = 0; g < 4; ++g) gxi[g] = gp[g * 256];
        }
      }
    }
    mbar_wait(bar0 + 8u * (uint32_t)cur, (uint32_t)(s >> 1) & 1u);        // all of h_s has landed in hbuf[cur]
    // partial W_hh h: my 4 rows x my 16 k, every utterance of the group.  v[r * 4 + u]
    float v[16];
#pragma unroll
    for (int u = 0; u < LSTM_UB; ++u) {
      const float4* hp = reinterpret_cast<const float4*>(&hbuf[cur][u][0]) + ks;
      float h[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 h4 = hp[i * 16];
        h[4 * i] = h4.x; h[4 * i + 1] = h4.y; h[4 * i + 2] = h4.z; h[4 * i + 3] = h4.w;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i += 2) fma2(a0, a1, w[r][i], w[r][i + 1], h[i], h[i + 1]);      // the same two chains, half the instructions
        v[r * 4 + u] = a0 + a1;
      }
    }
    // reduce-scatter over the 16 k-slices (lanes ks = 0..15 of each half warp): lane ks ends with the total of value ks
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
      const bool up = (ks & half) != 0;
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const float keep = up ? v[j + half] : v[j];
        const float send = up ? v[j] : v[j + half];
        v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
      }
    }
    gates_s[ks & 3][4 * rg + (ks >> 2)] = v[0];
    __syncthreads();
    if (cell && active) {
      const float ig = sigmoidf_acc(gates_s[cu][cj] + gxi[0]);
      const float fg = sigmoidf_acc(gates_s[cu][32 + cj] + gxi[1]);
      const float gg = tanhf(gates_s[cu][64 + cj] + gxi[2]);
      const float og = sigmoidf_acc(gates_s[cu][96 + cj] + gxi[3]);
      c = fg * c + ig * gg;
      const float hn = og * tanhf(c);
      const uint32_t laddr = hb0 + 4096u * (uint32_t)(cur ^ 1) + (uint32_t)(cu * 256 + hidx) * 4u;
      const uint32_t lbar = bar0 + 8u * (uint32_t)(cur ^ 1);
#pragma unroll
      for (int pcta = 0; pcta < NCTA; ++pcta) st_async_f32(map_to_peer(laddr, (uint32_t)pcta), hn, map_to_peer(lbar, (uint32_t)pcta));
      if (!is_pad) y[(size_t)(c_off + t) * ldy + dir * 256 + hcol] = hn;
    }
    // No barrier here.  gates_s is safe: nobody passes the wait for h_{s+1} before this CTA's cell threads have pushed, and they
    // push after reading gates_s.  hbuf[cur] is safe: see the header (a peer needs all of h_{s+1} before it can push h_{s+2}).
  }
  // A CTA must not retire while a peer may still push into it: wait for my own last h (then every push TO me has landed), and
  // meet the others (then every push FROM me has landed at its receiver, who waited the same way).
  mbar_wait(bar0 + 8u * (uint32_t)(maxsteps & 1), (uint32_t)(maxsteps >> 1) & 1u);
  cluster_arrive();
  cluster_wait();
}

}  // namespace

cudaError_t launch_lstm_cluster(const float* gx, const float* whh, const float* bias, float* y, int ldy, const LstmGroup* groups, int ngroups,
                                cudaStream_t s) {
  if (ngroups <= 0) return cudaSuccess;
  dim3 grid(NCTA, ngroups, 2);
  lstm_cluster_kernel<<<grid, NT, 0, s>>>(gx, whh, bias, y, ldy, groups);
  return launched();
}

}  // namespace ttsb
