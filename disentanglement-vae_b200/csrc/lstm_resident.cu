// Grid-resident LSTM recurrence for H = 512 / 1024 (BASELINE cfg 4): ONE launch runs all T steps of one direction with
// W_hh resident in TENSOR MEMORY across the whole grid.
//
// The cluster-resident kernels (lstm_tc.cu) stop at H = 256: at H = 1024 one direction's W_hh is 16 MB of fp16 (hi, lo)
// planes, more than an 8-CTA cluster holds.  Here the 4H x H matrix is tiled over the GRID:
//   CTA (ks, j), ks < KS = H / 256, j < H / 32:  gate rows {g * H + 32 j + u : g < 4, u < 32} (M = 128) x k in [256 ks, 256 ks + 256)
//   = 128 x 256 weights = 2 x 128 TMEM columns (hi | lo planes), next to two fp32 accumulators of 128 columns (N = batch <= 128):
//   all 512 columns of the SM's tensor memory.  H = 1024: 4 x 32 = 128 CTAs, one per SM.
// One step:
//   1. the 128 batch rows of h_{t-1}, k-slice ks, arrive as eight 16 KB bulk copies of tile-blocked fp16 planes (tc_planes.cuh)
//      that the CTAs themselves wrote to global memory (L2) at the end of the previous step;
//   2. 48 tcgen05.mma.kind::f16 (A = weights from tensor memory, B = state tile from shared memory, N = 128) accumulate the
//      k-slice's part of the recurrent product;
//   3. the KS CTAs of a cluster (the k-slices of one row block) reduce-scatter their partial sums over the batch columns with
//      bulk copies shared::cta -> shared::cluster that complete on the receiver's mbarrier (as lstm_tc.cu's backward);
//   4. each CTA finishes 128 / KS batch columns x 32 units: adds the hoisted input projection, applies the gates, updates c / h
//      (registers across all steps), writes gates / c / h for the backward pass and its 32 units of h_t into the OTHER
//      plane buffer;
//   5. grid-wide release / acquire on a step counter in global memory (every CTA resident: the grid is at most the SM count).
// The plane path (lstm_planes.cu: two launches per step, weights re-read from L2) remains for shapes this kernel does not
// take and for the backward pass.  Reference semantics: nn.LSTM inside vae/model.py:88-101, :152-165.
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "lstm_persist.cuh"
#include "tc_gemm.cuh"
#include "tc_gemm16.cuh"
#include "tc_planes.cuh"

namespace dvae {
namespace {

using namespace tc;

constexpr int kRThreads = 384;        // warps 0-3: control (0 = operand loader, 1 = MMA issue), warps 4-11: 256 workers
constexpr uint32_t kT_D1 = 0, kT_D2 = 128, kT_AHI = 256, kT_ALO = 384;
constexpr int kS_B = 0;               // state operand: 8 k-blocks x 16 KB [hi 8 KB | lo 8 KB]; after the MMAs: outgoing partial tiles
constexpr int kS_IN = 8 * 16384;      // incoming partial tiles of the KS - 1 peers: at most 3 x 16 KB / 1 x 32 KB
constexpr int kS_BAR = kS_IN + 49152;
constexpr int kS_BYTES = kS_BAR + 256;
constexpr float kLo = 2048.f, kLoInv = 1.f / 2048.f;

struct ResidentArgs {
  const float* w_hh;             // [4H, H] of this direction
  float* gates;                  // [T,B,4H] slab of this direction: input projection in, post-activation gates out
  float* cs;                     // [T,B,H] slab of this direction
  float* hs; int64_t ldhs;       // [T,B,*] layer output, this direction's column block
  const float* h0; const float* c0; int64_t ld0;
  float* hn; float* cn; int64_t ldn;
  const int64_t* lengths;
  uint8_t* planes;               // 2 x [128, H] state planes (zero-initialised: rows >= B stay zero), then the step counter
  unsigned* counter;
  int T, B, H, reverse;
};

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster), "r"(src_cta),
               "r"(bytes), "r"(mbar_cluster)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void add_release(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void split2(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * kLo);
}
__device__ __forceinline__ uint32_t pk(__half a, __half b) { return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16); }
__device__ __forceinline__ void stg_h(uint8_t* p, __half v) { *reinterpret_cast<__half*>(p) = v; }

template <int KS>
__global__ void __launch_bounds__(kRThreads, 1) lstm_resident_fwd_kernel(ResidentArgs p) {
  constexpr int NC = 128 / KS;                    // batch columns this CTA finishes
  constexpr int NP = NC / 8;                      // (unit, column) pairs per worker thread
  constexpr uint32_t kTile = 128 * NC * 4;        // one partial tile [NC cols][128 rows] fp32
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kS_BAR);
  uint64_t* full = bars;                          // [8] state k-block landed
  uint64_t* mma_done = bars + 8;
  uint64_t* red_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H, B = p.B, T = p.T;
  const int ks = (int)cluster_ctarank(), j = blockIdx.y;      // grid (KS, H / 32), cluster (KS, 1, 1)
  const unsigned n_ctas = gridDim.x * gridDim.y;
  const uint32_t sbase = smem_u32(smem);
  const int KBG = H / 32;                                      // k-blocks of a full state row

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1);
    mbar_init(mma_done, 1);
    mbar_init(red_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_barrier();           // every CTA's mbarriers exist before a peer can signal them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // worker roles.  read-out: TMEM lane quarter q (= gate q of unit 32 j + lane), batch columns [64 half, 64 half + 64).
  //               cell update: unit 32 j + lane, columns ks * NC + wrow + 8 i -- c / h live in registers across all steps.
  const int wk = tid - 128, q = warp & 3, half = (warp - 4) >> 2, wrow = wk >> 5;
  if (warp >= 4) {
    // resident weights: this thread owns gate row (q, lane) and k in [256 ks + 128 half, + 128): 2 x 64 TMEM columns
    const float* wrow_p = p.w_hh + (int64_t)(q * H + 32 * j + lane) * H + 256 * ks + 128 * half;
    const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + 64 * half;
#pragma unroll
    for (int part = 0; part < 4; ++part) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 x = *reinterpret_cast<const float4*>(wrow_p + part * 32 + 4 * e);
        __half h0, l0, h1, l1, h2, l2, h3, l3;
        split2(x.x, h0, l0); split2(x.y, h1, l1); split2(x.z, h2, l2); split2(x.w, h3, l3);
        hi[2 * e] = pk(h0, h1); hi[2 * e + 1] = pk(h2, h3);
        lo[2 * e] = pk(l0, l1); lo[2 * e + 1] = pk(l2, l3);
      }
      tmem_st16(tl + kT_AHI + part * 16, hi);
      tmem_st16(tl + kT_ALO + part * 16, lo);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  float c_reg[NP], h_reg[NP];
  int len_reg[NP];
  uint8_t* const planes0 = p.planes;
  const int64_t plane_bytes = (int64_t)KBG * 16384;            // one [128, H] buffer
  const int unit = 32 * j + lane;
  if (warp >= 4) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int b = ks * NC + wrow + 8 * i;
      c_reg[i] = (p.c0 && b < B) ? p.c0[(int64_t)b * p.ld0 + unit] : 0.f;
      h_reg[i] = (p.h0 && b < B) ? p.h0[(int64_t)b * p.ld0 + unit] : 0.f;
      len_reg[i] = b < B ? (p.lengths ? (int)p.lengths[b] : T) : 0;
    }
  }
  // operand planes of this CTA's 32 units for its NC columns: tile j of buffer `buf`
  auto publish = [&](int buf) {
    uint8_t* tile = planes0 + buf * plane_bytes + (int64_t)j * 16384;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int b = ks * NC + wrow + 8 * i;
      if (b < B) {
        __half hi, lo;
        split2(h_reg[i], hi, lo);
        const int off = (b >> 3) * 512 + (lane >> 3) * 128 + (b & 7) * 16 + (lane & 7) * 2;
        stg_h(tile + off, hi);
        stg_h(tile + 8192 + off, lo);
      }
    }
    fence_proxy_async_all();         // generic-proxy global stores -> visible to the consumers' bulk copies (async proxy)
    workers_sync();
    if (wk == 0) {
      __threadfence();
      add_release(p.counter, 1u);
    }
  };
  unsigned pubs = 0;                 // states published so far by every CTA (the counter reaches n_ctas * pubs)
  if (p.h0) {
    if (warp >= 4) publish(1);       // the initial state plays "h_{-1}": buffer (-1) & 1
    pubs = 1;
  }

  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  int ms = 0;                        // steps that ran the recurrent product (mbarrier phases)
  for (int s = 0; s < T; ++s) {
    const int t = p.reverse ? T - 1 - s : s;
    const bool has_prev = s > 0 || p.h0 != nullptr;
    const uint32_t ph = ms & 1;
    if (warp == 0) {
      // ===== operand loader =====
      if (lane == 0 && has_prev) {
        const unsigned target = n_ctas * pubs;
        unsigned spins = 0;
        while (ld_acquire(p.counter) < target) {
          if (++spins > (1u << 23)) {
            printf("dvae lstm_resident: grid barrier timed out (block %d,%d step %d)\n", blockIdx.x, blockIdx.y, s);
            __trap();
          }
        }
        fence_proxy_async_all();
        const uint8_t* src = planes0 + ((s + 1) & 1) * plane_bytes + (int64_t)(8 * ks) * 16384;      // h_{s-1}: buffer (s - 1) & 1
#pragma unroll
        for (int kb = 0; kb < 8; ++kb) {
          mbar_expect_tx(&full[kb], 16384);
          bulk_g2s(sbase + kS_B + kb * 16384, src + kb * 16384, 16384, &full[kb]);
        }
      }
    } else if (warp == 1) {
      // ===== MMA issue: partial[128 gate rows, 128 batch] over this k-slice =====
      if (has_prev && elect_one()) {
        constexpr uint32_t d_hi = ((512 >> 4) & 0x3FFF) | (1u << 14);          // SBO 512, descriptor version, no swizzle
        constexpr uint32_t d_lo = ((128 >> 4) & 0x3FFF) << 16;                  // LBO 128
        auto desc = [&](uint32_t addr16) { return ((uint64_t)d_hi << 32) | (uint64_t)(d_lo | (addr16 & 0x3FFF)); };
#pragma unroll 1
        for (int kb = 0; kb < 8; ++kb) {
          mbar_wait(&full[kb], ph);
          tc_fence_after();
          const uint32_t b16 = (sbase + kS_B + kb * 16384) >> 4;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t ahi = tmem + kT_AHI + kb * 16 + k * 8, alo = tmem + kT_ALO + kb * 16 + k * 8;
            const uint32_t bh = b16 + k * 16, bl = b16 + 512 + k * 16;          // lo plane 8 KB later; 16 k = two 128-byte core matrices
            mma_f16_ts(tmem + kT_D1, ahi, desc(bh), idesc, (kb | k) ? 1u : 0u);
            mma_f16_ts(tmem + kT_D2, ahi, desc(bl), idesc, (kb | k) ? 1u : 0u);
            mma_f16_ts(tmem + kT_D2, alo, desc(bh), idesc, 1u);
          }
        }
        tc_commit(mma_done);
      }
      __syncwarp();
    } else if (warp >= 4) {
      // ===== workers =====
      // hoisted input projection of this thread's (unit, column) pairs, fetched before the recurrent product is waited for
      float gx[NP][4];
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int b = ks * NC + wrow + 8 * i;
        const float* g = p.gates + ((int64_t)t * B + b) * 4 * H + unit;
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) gx[i][gt] = b < B ? g[(int64_t)gt * H] : 0.f;
      }
      float* own = reinterpret_cast<float*>(smem + kS_B);            // outgoing tiles [dest][col][row]; dest == ks stays here
      const float* inc = reinterpret_cast<const float*>(smem + kS_IN);
      if (has_prev) {
        if (wk == 0) mbar_expect_tx(red_full, (KS - 1) * kTile);
        if (lane == 0 && warp == 4) mbar_wait(mma_done, ph);
        workers_sync();
        tc_fence_after();
        // accumulators -> partial tiles (column-major inside a tile: the 32 lanes of a warp write 32 consecutive rows)
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + 64 * half;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v[16], w[16];
          tmem_ld16_pair(ta + kT_D1 + 16 * c, ta + kT_D2 + 16 * c, v, w);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = 64 * half + 16 * c + i;
            own[(col / NC) * (128 * NC) + (col % NC) * 128 + q * 32 + lane] = fmaf(w[i], kLoInv, v[i]);
          }
        }
        tc_fence_before();
        fence_proxy_async();           // shared-memory stores -> visible to the bulk-copy engine
        workers_sync();
        if (warp == 4 && lane < KS - 1) {
          const int peer = lane + (lane >= ks ? 1 : 0);
          const uint32_t src = sbase + kS_B + peer * kTile;
          const uint32_t slot = ks < peer ? ks : ks - 1;             // my tile's slot in the peer's incoming area
          bulk_s2peer(mapa(sbase + kS_IN + slot * kTile, peer), src, kTile, mapa(smem_u32(red_full), peer));
        }
        if (lane == 0 && warp == 4) mbar_wait(red_full, ph);
        workers_sync();
      }
      // cell update
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const int bc = wrow + 8 * i, b = ks * NC + bc;
        float pre[4];
#pragma unroll
        for (int gt = 0; gt < 4; ++gt) {
          float a = gx[i][gt];
          if (has_prev) {
            a += own[ks * (128 * NC) + bc * 128 + gt * 32 + lane];
#pragma unroll
            for (int r = 0; r < KS - 1; ++r) a += inc[r * (128 * NC) + bc * 128 + gt * 32 + lane];
          }
          pre[gt] = a;
        }
        if (b < B) {
          float* g = p.gates + ((int64_t)t * B + b) * 4 * H + unit;
          float out = 0.f;
          if (t < len_reg[i]) {
            const float ig = sigmoidf_(pre[0]), fg = sigmoidf_(pre[1]), gg = tanhf(pre[2]), og = sigmoidf_(pre[3]);
            c_reg[i] = fmaf(fg, c_reg[i], ig * gg);
            h_reg[i] = og * tanhf(c_reg[i]);
            out = h_reg[i];
            g[0] = ig; g[(int64_t)H] = fg; g[(int64_t)2 * H] = gg; g[(int64_t)3 * H] = og;
          } else {            // frozen / padded row: zero gates and output, carried state
            g[0] = 0.f; g[(int64_t)H] = 0.f; g[(int64_t)2 * H] = 0.f; g[(int64_t)3 * H] = 0.f;
          }
          p.cs[((int64_t)t * B + b) * H + unit] = c_reg[i];
          p.hs[((int64_t)t * B + b) * p.ldhs + unit] = out;
        }
      }
      if (s + 1 < T) publish(s & 1);          // h_s -> buffer s & 1
    }
    if (has_prev) ++ms;
    if (s + 1 < T) ++pubs;
  }
  if (warp >= 4) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int b = ks * NC + wrow + 8 * i;
      if (b < B) {
        if (p.hn) p.hn[(int64_t)b * p.ldn + unit] = h_reg[i];
        if (p.cn) p.cn[(int64_t)b * p.ldn + unit] = c_reg[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier();           // no CTA leaves while a peer's bulk copy may still target it
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int KS>
int launch_resident(const ResidentArgs& a, cudaStream_t st) {
  static bool ready = false;
  if (!ready) {
    DVAE_CUDA(cudaFuncSetAttribute(lstm_resident_fwd_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kS_BYTES));
    ready = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(KS, a.H / 32, 1);
  cfg.blockDim = dim3(kRThreads);
  cfg.dynamicSmemBytes = kS_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = KS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DVAE_CUDA(cudaLaunchKernelEx(&cfg, lstm_resident_fwd_kernel<KS>, a));
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

}  // namespace

// H = 512 (32 CTAs) or 1024 (128 CTAs), at most 128 batch rows; DVAE_LSTM_IMPL=planes keeps the per-step plane path (A/B tests)
bool resident_lstm_supported(int B, int H) {
  const char* e = getenv("DVAE_LSTM_IMPL");
  if (e && !strcmp(e, "planes")) return false;
  return (H == 512 || H == 1024) && B >= 1 && B <= 128;
}

int64_t resident_lstm_ws_floats(int H) { return 2 * tc16::plane_floats(128, H) + 4; }

// forward recurrence of one layer (gates hold the input projection); `pws`: resident_lstm_ws_floats(H) floats, 16-byte aligned
int resident_lstm_fwd(int T, int B, int H, int D, const float* const* w_hh, const float* h0, const float* c0, int64_t ld0,
                      int64_t dir0, const int64_t* lengths, float* hs, int64_t ldhs, float* hn, float* cn, int64_t ldn,
                      int64_t dirn, float* gates, float* cs, float* pws, cudaStream_t st) {
  const int64_t slab = (int64_t)T * B * 4 * H, cslab = (int64_t)T * B * H;
  const int64_t nfl = resident_lstm_ws_floats(H);
  for (int d = 0; d < D; ++d) {       // a direction occupies the whole grid (128 SMs at H = 1024): one after the other
    DVAE_CUDA(cudaMemsetAsync(pws, 0, sizeof(float) * nfl, st));
    ResidentArgs a;
    a.w_hh = w_hh[d]; a.gates = gates + d * slab; a.cs = cs + d * cslab; a.hs = hs + d * H; a.ldhs = ldhs;
    a.h0 = h0 ? h0 + d * dir0 : nullptr; a.c0 = c0 ? c0 + d * dir0 : nullptr; a.ld0 = ld0;
    a.hn = hn ? hn + d * dirn : nullptr; a.cn = cn ? cn + d * dirn : nullptr; a.ldn = ldn;
    a.lengths = lengths; a.planes = reinterpret_cast<uint8_t*>(pws); a.counter = reinterpret_cast<unsigned*>(pws + nfl - 4);
    a.T = T; a.B = B; a.H = H; a.reverse = d;
    int rc = H == 1024 ? launch_resident<4>(a, st) : launch_resident<2>(a, st);
    if (rc) return rc;
  }
  return DVAE_OK;
}

}  // namespace dvae
