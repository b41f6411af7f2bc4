// Persistent LSTM recurrence for H in {64,128,256}: ONE launch runs all T time steps.
//
// A thread-block cluster of 8 CTAs owns (direction, 16 batch rows).  CTA `rank` keeps the
// recurrent weights of its H/8 hidden units (all four gates: a [H/2, H] slice of W_hh, 128 KB at
// H = 256) resident in shared memory for the whole sequence, so W_hh is read from HBM/L2 exactly
// once per layer call instead of once per step.  Each step:
//   1. every CTA stages the 16 carried h rows (16 x H fp32, L2-resident exchange buffer) in SMEM;
//   2. a K-split, register-tiled mini-GEMM (8 rows x 2 gate-columns per thread, float4 shared
//      loads, conflict-free row padding) produces the recurrent pre-activations of its slice;
//   3. H/2 epilogue threads add the hoisted input projection, apply the gate non-linearities,
//      update c (kept in registers across steps) and h, store gates / c / h for backward;
//   4. the new h slice is published and the cluster meets at the hardware cluster barrier.
// The backward kernel is the mirror image with the transposed slice W_hh[:, units] resident and
// the gate gradients dG[t] as the exchanged quantity; carried dh / dc live in registers.
// 16 clusters x 8 CTAs = 128 of the 148 SMs at the cfg-2 shape (B = 128, two directions).
#include <cooperative_groups.h>

#include "common.cuh"
#include "lstm_persist.cuh"

namespace cg = cooperative_groups;

namespace dvae {

constexpr int kPRows = 16;        // batch rows per cluster
constexpr int kPThreads = 256;
constexpr int kClusterSize = 8;
constexpr int kMaxResidentClusters = 15;

// acc[8][2] = sum over this thread's K range of A[rh*8 + r][k] * W[col(c)][k]
//   A: smem [16][K + 4];  W: smem [NCOLS][K + 4];  thread tile = 8 rows x cols {cp, cp + NCOLS/2}
template <int NCOLS, int K, int KPS>
__device__ __forceinline__ void persist_gemm(const float* __restrict__ a_s, const float* __restrict__ w_s,
                                             float* __restrict__ part_s) {
  constexpr int LD = K + 4;
  const int tid = threadIdx.x;
  const int tile = tid % NCOLS, ks = tid / NCOLS;
  const int rh = tile / (NCOLS / 2), cp = tile % (NCOLS / 2);
  const float* a_base = a_s + (rh * 8) * LD + ks * KPS;
  const float* w0 = w_s + cp * LD + ks * KPS;
  const float* w1 = w_s + (cp + NCOLS / 2) * LD + ks * KPS;
  // packed fp32 FMA (FFMA2): each accumulator is an (even k, odd k) pair, so both operands of every
  // FFMA2 are the natural register pairs of the LDS.128 results -- no duplication moves.  Scalar
  // FFMA issues at half rate on sm_100; FFMA2 is what reaches the fp32 peak.
  float2 acc[8][2];
#pragma unroll
  for (int r = 0; r < 8; ++r) { acc[r][0] = make_float2(0.f, 0.f); acc[r][1] = make_float2(0.f, 0.f); }
#pragma unroll 2
  for (int kq = 0; kq < KPS / 4; ++kq) {
    const float4 b0 = *reinterpret_cast<const float4*>(w0 + kq * 4);
    const float4 b1 = *reinterpret_cast<const float4*>(w1 + kq * 4);
    const float2 b0l = make_float2(b0.x, b0.y), b0h = make_float2(b0.z, b0.w);
    const float2 b1l = make_float2(b1.x, b1.y), b1h = make_float2(b1.z, b1.w);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(a_base + r * LD + kq * 4);
      const float2 al = make_float2(a.x, a.y), ah = make_float2(a.z, a.w);
      acc[r][0] = __ffma2_rn(al, b0l, acc[r][0]); acc[r][0] = __ffma2_rn(ah, b0h, acc[r][0]);
      acc[r][1] = __ffma2_rn(al, b1l, acc[r][1]); acc[r][1] = __ffma2_rn(ah, b1h, acc[r][1]);
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    part_s[(ks * kPRows + rh * 8 + r) * NCOLS + cp] = acc[r][0].x + acc[r][0].y;
    part_s[(ks * kPRows + rh * 8 + r) * NCOLS + cp + NCOLS / 2] = acc[r][1].x + acc[r][1].y;
  }
}

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define LSTM_MARK(i) do { if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0 && s == 6) p.dbg[i] = gtime_ns(); } while (0)

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }


template <int H>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kPThreads, 1)
lstm_persist_fwd_kernel(PersistFwdArgs p) {
  constexpr int UPC = H / kClusterSize;       // hidden units per CTA
  constexpr int NCOLS = 4 * UPC;              // gate columns per CTA
  constexpr int SPLITS = kPThreads / NCOLS;   // K splits
  constexpr int KPS = H / SPLITS;
  constexpr int LD = H + 4;
  constexpr int NEPI = kPRows * UPC / 4;      // epilogue threads (one float4 of units each)
  static_assert(KPS % 4 == 0 && SPLITS >= 1 && NEPI <= kPThreads, "unsupported H");
  extern __shared__ __align__(16) float sm[];
  float* w_s = sm;                             // [NCOLS][LD]
  float* a_s = w_s + NCOLS * LD;               // [16][LD]
  float* part_s = a_s + kPRows * LD;           // [SPLITS][16][NCOLS]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kClusterSize;
  const int d = p.d_off + cid / p.n_slices, slice = cid % p.n_slices;
  const int b0 = slice * kPRows, u0 = rank * UPC;
  const int tid = threadIdx.x, B = p.B, T = p.T;

  // resident weight slice: local row j = g*UPC + u  <-  W_hh[g*H + u0 + u][:]
  {
    const float* W = p.w_hh[d];
    for (int v = tid; v < NCOLS * (H / 4); v += kPThreads) {
      int j = v / (H / 4), kq = v % (H / 4);
      int g = j / UPC, u = j % UPC;
      *reinterpret_cast<float4*>(&w_s[j * LD + kq * 4]) =
          *reinterpret_cast<const float4*>(W + (int64_t)(g * H + u0 + u) * H + kq * 4);
    }
  }
  // epilogue role: row er, units u0 + euq*4 .. +3
  const bool epi = tid < NEPI;
  const int er = tid / (UPC / 4), euq = tid % (UPC / 4);
  const int eb = b0 + er, eu = u0 + euq * 4;
  const bool erow = epi && eb < B;
  float c_reg[4] = {0.f, 0.f, 0.f, 0.f}, h_reg[4] = {0.f, 0.f, 0.f, 0.f};
  int64_t len = T;
  if (erow) {
    if (p.lengths) len = p.lengths[eb];
    if (p.c0) {
      const float4 c4 = *reinterpret_cast<const float4*>(p.c0 + d * p.dir0 + (int64_t)eb * p.ld0 + eu);
      c_reg[0] = c4.x; c_reg[1] = c4.y; c_reg[2] = c4.z; c_reg[3] = c4.w;
    }
    if (p.h0) {
      const float4 h4 = *reinterpret_cast<const float4*>(p.h0 + d * p.dir0 + (int64_t)eb * p.ld0 + eu);
      h_reg[0] = h4.x; h_reg[1] = h4.y; h_reg[2] = h4.z; h_reg[3] = h4.w;
    }
  }
  const int64_t state_stride = (int64_t)p.D * B * H;

  for (int s = 0; s < T; ++s) {
    const int t = d == 0 ? s : T - 1 - s;
    LSTM_MARK(0);
    // (1) prefetch the hoisted input projection for this thread's 4 units x 4 gates
    float4 pre[4];
    const int64_t gi = (((int64_t)d * T + t) * B + eb) * 4 * H + eu;
    if (erow) {
#pragma unroll
      for (int g = 0; g < 4; ++g) pre[g] = *reinterpret_cast<const float4*>(p.gates + gi + g * H);
    }
    // (2) stage the 16 carried h rows: all loads in flight before the first shared store
    {
      constexpr int NV = kPRows * (H / 4) / kPThreads;
      float4 val[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = tid + i * kPThreads, r = v / (H / 4), kq = v % (H / 4), b = b0 + r;
        val[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < B) {
          if (s == 0) {
            if (p.h0) val[i] = *reinterpret_cast<const float4*>(p.h0 + d * p.dir0 + (int64_t)b * p.ld0 + kq * 4);
          } else {
            val[i] = ldcg4(p.hstate + (s & 1) * state_stride + ((int64_t)d * B + b) * H + kq * 4);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int v = tid + i * kPThreads, r = v / (H / 4), kq = v % (H / 4);
        *reinterpret_cast<float4*>(&a_s[r * LD + kq * 4]) = val[i];
      }
    }
    __syncthreads();
    LSTM_MARK(1);
    // (3) recurrent pre-activations of this CTA's gate columns
    persist_gemm<NCOLS, H, KPS>(a_s, w_s, part_s);
    __syncthreads();
    LSTM_MARK(2);
    // (4) gates, cell update, outputs
    if (erow) {
      float gate[4][4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float pg[4] = {pre[g].x, pre[g].y, pre[g].z, pre[g].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = pg[i];
#pragma unroll
          for (int k = 0; k < SPLITS; ++k) a += part_s[(k * kPRows + er) * NCOLS + g * UPC + euq * 4 + i];
          gate[g][i] = a;
        }
      }
      const bool live = t < len;
      float hs_out[4];
      if (live) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ig = sigmoidf_(gate[0][i]), fg = sigmoidf_(gate[1][i]);
          const float gg = tanhf(gate[2][i]), og = sigmoidf_(gate[3][i]);
          c_reg[i] = fmaf(fg, c_reg[i], ig * gg);
          h_reg[i] = og * tanhf(c_reg[i]);
          gate[0][i] = ig; gate[1][i] = fg; gate[2][i] = gg; gate[3][i] = og;
          hs_out[i] = h_reg[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          gate[0][i] = gate[1][i] = gate[2][i] = gate[3][i] = 0.f;
          hs_out[i] = 0.f;
        }
      }
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<float4*>(p.gates + gi + g * H) = make_float4(gate[g][0], gate[g][1], gate[g][2], gate[g][3]);
      *reinterpret_cast<float4*>(p.cs + (((int64_t)d * T + t) * B + eb) * H + eu) =
          make_float4(c_reg[0], c_reg[1], c_reg[2], c_reg[3]);
      *reinterpret_cast<float4*>(p.hs + ((int64_t)t * B + eb) * p.ldhs + d * H + eu) =
          make_float4(hs_out[0], hs_out[1], hs_out[2], hs_out[3]);
      *reinterpret_cast<float4*>(p.hstate + ((s + 1) & 1) * state_stride + ((int64_t)d * B + eb) * H + eu) =
          make_float4(h_reg[0], h_reg[1], h_reg[2], h_reg[3]);
    }
    LSTM_MARK(3);
    cluster.sync();      // barrier.cluster arrive.release / wait.acquire orders the global stores above
    LSTM_MARK(4);
  }
  if (erow) {
    if (p.hn) *reinterpret_cast<float4*>(p.hn + d * p.dirn + (int64_t)eb * p.ldn + eu) = make_float4(h_reg[0], h_reg[1], h_reg[2], h_reg[3]);
    if (p.cn) *reinterpret_cast<float4*>(p.cn + d * p.dirn + (int64_t)eb * p.ldn + eu) = make_float4(c_reg[0], c_reg[1], c_reg[2], c_reg[3]);
  }
}


template <int H>
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kPThreads, 1)
lstm_persist_bwd_kernel(PersistBwdArgs p) {
  constexpr int UPC = H / kClusterSize;
  constexpr int NCOLS = UPC;                   // output columns = this CTA's hidden units
  constexpr int K = 4 * H;
  constexpr int SPLITS = kPThreads / NCOLS;
  constexpr int KPS = K / SPLITS;
  constexpr int LD = K + 4;
  constexpr int NEPI = kPRows * UPC / 4;
  static_assert(KPS % 4 == 0 && NCOLS % 2 == 0 && NEPI <= kPThreads, "unsupported H");
  extern __shared__ __align__(16) float sm[];
  float* w_s = sm;                             // [UPC][LD]: w_s[u][j] = W_hh[j][u0 + u]
  float* a_s = w_s + NCOLS * LD;               // [16][LD] staged dG rows
  float* part_s = a_s + kPRows * LD;           // [SPLITS][16][NCOLS]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kClusterSize;
  const int d = p.d_off + cid / p.n_slices, slice = cid % p.n_slices;
  const int b0 = slice * kPRows, u0 = rank * UPC;
  const int tid = threadIdx.x, B = p.B, T = p.T;
  {
    const float* W = p.w_hh[d];
    for (int v = tid; v < K * UPC; v += kPThreads) {
      const int j = v / UPC, u = v % UPC;      // consecutive threads read consecutive units of one W row
      w_s[u * LD + j] = W[(int64_t)j * H + u0 + u];
    }
  }
  const bool epi = tid < NEPI;
  const int er = tid / (UPC / 4), euq = tid % (UPC / 4);
  const int eb = b0 + er, eu = u0 + euq * 4;
  const bool erow = epi && eb < B;
  float carry[4] = {0.f, 0.f, 0.f, 0.f}, dc[4] = {0.f, 0.f, 0.f, 0.f};
  int64_t len = T;
  if (erow) {
    if (p.lengths) len = p.lengths[eb];
    if (p.d_hn) { const float4 v = *reinterpret_cast<const float4*>(p.d_hn + d * p.dirn + (int64_t)eb * p.ldn + eu); carry[0] = v.x; carry[1] = v.y; carry[2] = v.z; carry[3] = v.w; }
    if (p.d_cn) { const float4 v = *reinterpret_cast<const float4*>(p.d_cn + d * p.dirn + (int64_t)eb * p.ldn + eu); dc[0] = v.x; dc[1] = v.y; dc[2] = v.z; dc[3] = v.w; }
  }
  const int nsteps = T + ((p.d_h0 || p.d_c0) ? 1 : 0);
  for (int s = 0; s < nsteps; ++s) {
    const bool final_ = (s == T);
    const int t = d == 0 ? T - 1 - s : s;
    const int t_next = d == 0 ? t + 1 : t - 1;
    const int t_prev = d == 0 ? t - 1 : t + 1;
    // prefetch the epilogue operands
    float4 g4[4], c4 = make_float4(0.f, 0.f, 0.f, 0.f), cp4 = make_float4(0.f, 0.f, 0.f, 0.f), dh4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t gi = (((int64_t)d * T + t) * B + eb) * 4 * H + eu;
    const bool live = erow && !final_ && t < len;
    if (live) {
#pragma unroll
      for (int g = 0; g < 4; ++g) g4[g] = *reinterpret_cast<const float4*>(p.gates + gi + g * H);
      c4 = *reinterpret_cast<const float4*>(p.cs + (((int64_t)d * T + t) * B + eb) * H + eu);
      if (t_prev >= 0 && t_prev < T) cp4 = *reinterpret_cast<const float4*>(p.cs + (((int64_t)d * T + t_prev) * B + eb) * H + eu);
      else if (p.c0) cp4 = *reinterpret_cast<const float4*>(p.c0 + d * p.dir0 + (int64_t)eb * p.ld0 + eu);
      if (p.d_hs) dh4 = *reinterpret_cast<const float4*>(p.d_hs + ((int64_t)t * B + eb) * p.lddhs + d * H + eu);
    }
    float rec[4] = {0.f, 0.f, 0.f, 0.f};
    if (s > 0) {
      const float* src = p.gates + ((int64_t)d * T + t_next) * B * K;
      constexpr int NV = kPRows * (K / 4) / kPThreads;      // 16 at H = 256: two batches of 8 loads
      constexpr int NB = NV > 8 ? 8 : NV;
#pragma unroll
      for (int i0 = 0; i0 < NV; i0 += NB) {
        float4 val[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          const int v = tid + (i0 + i) * kPThreads, r = v / (K / 4), kq = v % (K / 4), b = b0 + r;
          val[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (b < B) val[i] = ldcg4(src + (int64_t)b * K + kq * 4);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          const int v = tid + (i0 + i) * kPThreads, r = v / (K / 4), kq = v % (K / 4);
          *reinterpret_cast<float4*>(&a_s[r * LD + kq * 4]) = val[i];
        }
      }
      __syncthreads();
      persist_gemm<NCOLS, K, KPS>(a_s, w_s, part_s);
      __syncthreads();
      if (erow) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int k = 0; k < SPLITS; ++k) rec[i] += part_s[(k * kPRows + er) * NCOLS + euq * 4 + i];
      }
    }
    if (erow) {
      if (final_) {
        if (p.d_h0) *reinterpret_cast<float4*>(p.d_h0 + d * p.dird0 + (int64_t)eb * p.ldd0 + eu) =
            make_float4(carry[0] + rec[0], carry[1] + rec[1], carry[2] + rec[2], carry[3] + rec[3]);
        if (p.d_c0) *reinterpret_cast<float4*>(p.d_c0 + d * p.dird0 + (int64_t)eb * p.ldd0 + eu) = make_float4(dc[0], dc[1], dc[2], dc[3]);
      } else if (!live) {
        // frozen row: pass the carried gradients through, emit zero gate gradients
#pragma unroll
        for (int i = 0; i < 4; ++i) carry[i] += rec[i];
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<float4*>(p.gates + gi + g * H) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        const float ig[4] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w}, fg[4] = {g4[1].x, g4[1].y, g4[1].z, g4[1].w};
        const float gg[4] = {g4[2].x, g4[2].y, g4[2].z, g4[2].w}, og[4] = {g4[3].x, g4[3].y, g4[3].z, g4[3].w};
        const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, cpv[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
        const float dhs[4] = {dh4.x, dh4.y, dh4.z, dh4.w};
        float o[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dh = carry[i] + rec[i] + dhs[i];
          const float tc = tanhf(cc[i]);
          const float dct = fmaf(dh * og[i], 1.f - tc * tc, dc[i]);
          o[0][i] = dct * gg[i] * ig[i] * (1.f - ig[i]);
          o[1][i] = dct * cpv[i] * fg[i] * (1.f - fg[i]);
          o[2][i] = dct * ig[i] * (1.f - gg[i] * gg[i]);
          o[3][i] = dh * tc * og[i] * (1.f - og[i]);
          carry[i] = 0.f;
          dc[i] = dct * fg[i];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<float4*>(p.gates + gi + g * H) = make_float4(o[g][0], o[g][1], o[g][2], o[g][3]);
      }
    }
    cluster.sync();      // barrier.cluster arrive.release / wait.acquire orders the global stores above
  }
}

// ---- host side ----------------------------------------------------------------------------------
template <int H>
static size_t persist_fwd_smem() { return sizeof(float) * ((size_t)(H / 2 + kPRows) * (H + 4) + 4096); }
template <int H>
static size_t persist_bwd_smem() { return sizeof(float) * ((size_t)(H / 8 + kPRows) * (4 * H + 4) + 4096); }

bool persist_supported(int B, int H, int D, const void* const* ptrs, int nptr, const int64_t* lds, int nld) {
  if (!(H == 64 || H == 128 || H == 256)) return false;
  for (int i = 0; i < nptr; ++i)
    if (ptrs[i] && (reinterpret_cast<uintptr_t>(ptrs[i]) & 15)) return false;
  for (int i = 0; i < nld; ++i)
    if (lds[i] % 4) return false;
  return B > 0 && (D == 1 || D == 2);
}

template <int H>
static int launch_fwd(const PersistFwdArgs& a, cudaStream_t st) {
  const size_t smem = persist_fwd_smem<H>();
  static bool ready = false;
  if (!ready) {
    DVAE_CUDA(cudaFuncSetAttribute(lstm_persist_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ready = true;
  }
  // at one CTA per SM a B200 keeps at most 15 clusters of 8 resident (measured with
  // cudaOccupancyMaxActiveClusters, profiles/probes/cluster_occupancy.cu): when both directions
  // together need more but one direction fits, run them back to back instead of spilling one
  // cluster into a second wave.
  const bool split = a.D == 2 && 2 * a.n_slices > kMaxResidentClusters && a.n_slices <= kMaxResidentClusters;
  for (int d0 = 0; d0 < (split ? 2 : 1); ++d0) {
    PersistFwdArgs b = a;
    b.d_off = d0;
    lstm_persist_fwd_kernel<H><<<kClusterSize * a.n_slices * (split ? 1 : a.D), kPThreads, smem, st>>>(b);
    DVAE_LAUNCH_CHECK();
  }
  return DVAE_OK;
}

template <int H>
static int launch_bwd(const PersistBwdArgs& a, cudaStream_t st) {
  const size_t smem = persist_bwd_smem<H>();
  static bool ready = false;
  if (!ready) {
    DVAE_CUDA(cudaFuncSetAttribute(lstm_persist_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ready = true;
  }
  const bool split = a.D == 2 && 2 * a.n_slices > kMaxResidentClusters && a.n_slices <= kMaxResidentClusters;
  for (int d0 = 0; d0 < (split ? 2 : 1); ++d0) {
    PersistBwdArgs b = a;
    b.d_off = d0;
    lstm_persist_bwd_kernel<H><<<kClusterSize * a.n_slices * (split ? 1 : a.D), kPThreads, smem, st>>>(b);
    DVAE_LAUNCH_CHECK();
  }
  return DVAE_OK;
}

int persist_fwd(int H, const PersistFwdArgs& a, cudaStream_t st) {
  if (tc_lstm_supported(H)) return tc_lstm_fwd(a, st);
  if (H == 256) return launch_fwd<256>(a, st);
  if (H == 128) return launch_fwd<128>(a, st);
  return launch_fwd<64>(a, st);
}

int persist_bwd(int H, const PersistBwdArgs& a, cudaStream_t st) {
  if (tc_lstm_supported(H)) return tc_lstm_bwd(a, st);
  if (H == 256) return launch_bwd<256>(a, st);
  if (H == 128) return launch_bwd<128>(a, st);
  return launch_bwd<64>(a, st);
}

}  // namespace dvae
