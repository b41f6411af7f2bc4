// Persistent LSTM recurrence on the 5th-generation tensor cores (H = 256): ONE launch runs all T steps.
//
// A cluster of 8 CTAs owns (direction, NB = 16 batch rows).  CTA `rank` owns hidden units [32*rank, 32*rank+32)
// and keeps their recurrent weights -- a [128 gate rows, 256] slice of W_hh -- resident in shared memory for the
// whole sequence as TWO fp16 planes (hi = fp16(w), lo = fp16((w - hi) * 2^11)), in the UMMA no-swizzle K-major
// core-matrix layout.  The carried state is split the same way, and each step is
//     D1  = A_hi * B_hi                      (tcgen05.mma.kind::f16, M = 128, N = 16, K = 16, fp32 accumulate in TMEM)
//     D2  = A_hi * B_lo + A_lo * B_hi        (second accumulator; the 2^11 factor keeps `lo` in fp16's normal range)
//     out = D1 + D2 * 2^-11                  -> 22 mantissa bits per operand: fp32-grade pre-activations
// i.e. 48 MMAs per step issued by one thread.  The epilogue reads the accumulators with tcgen05.ld (TMEM lane =
// gate row, column = batch row), adds the hoisted input projection, applies the gate non-linearities, exchanges the
// four gates of a unit through shared memory, updates c / h (registers across steps) and writes gates / c / h for
// the backward pass.
//
// State exchange without a cluster barrier: each of the 16 warps owns one batch row in the cell update; it writes the
// row's 32 new state values (fp16 hi / lo, already in operand layout) into its own CTA's operand buffer, stores the
// same eight 16-byte chunks into the 7 peers' buffers (st.shared::cluster), fences generic -> async proxy and arrives
// (release.cluster) on all 8 CTAs' "operand full" mbarriers.  The MMA warp waits (acquire.cluster) for 8 x 16 arrivals,
// double-buffered by step parity -- no barrier.cluster and no __syncthreads on the exchange path.
//
// Backward (lstm_tc_bwd_kernel) is K-split instead: the CTA keeps W_hh[own gate rows, :]^T (M = 256 units, K = 128)
// resident (in tensor memory), multiplies it with its OWN gate gradients dG[t+1] (per-row power-of-two scaling keeps them inside fp16
// range; exact, undone in the epilogue), and the partial d(h) values are reduce-scattered straight from registers into
// the unit owners' shared memory with the same remote-store + mbarrier mechanism.  Carried dh / dc stay in registers.
//
// Reference semantics: nn.LSTM inside vae/model.py:88-101 (packed, variable length) and :152-165 (decoder).
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "lstm_persist.cuh"
#include "tc_gemm.cuh"

namespace cg = cooperative_groups;

namespace dvae {
namespace {

using namespace tc;

constexpr int kH = 256;          // hidden size this kernel is built for
constexpr int kCS = 8;           // CTAs per cluster
constexpr int kUPC = kH / kCS;   // 32 hidden units per CTA
constexpr int kNB = 16;          // batch rows per cluster (MMA N)
constexpr int kGT = 512;         // threads per row group: 4 TMEM lane quadrants x 4 column groups; one warp per batch row in the cell update
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (lane = row, each 32-bit column = two consecutive k) stays in tensor
// memory, so an MMA reads only its small B tile from shared memory
// registers -> 32 lanes x 16 consecutive 32-bit columns
// Programmatic dependent launch: the kernel is launched while its stream predecessor (typically the hoisted input
// projection GEMM) is still running; everything up to pdl_wait() -- barrier init, TMEM allocation, the split of the
// resident weights into tensor memory -- overlaps the predecessor's tail, and nothing the predecessor wrote is read
// before it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// kind::f16, A = B = fp16, D = fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * kLoScale);
}
__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}
__device__ __forceinline__ void sts_h(uint32_t addr, __half v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(v)) : "memory");
}
// own shared memory -> a peer CTA's shared memory; completes (complete_tx) on the peer's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}
// hardware barrier of one 512-thread row group (barrier 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int grp, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(nthreads) : "memory"); }
// 32 lanes x 4 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// Gate non-linearities on the SFU (ex2.approx + rcp.approx): absolute error ~2e-7, i.e. fp32 rounding level for
// values in (-1, 1); the precise libm versions cost ~40 instructions each on a latency-exposed path.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.f, __fdividef(1.f, 1.f + __expf(-2.f * x)), -1.f); }
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// optional milestone timestamps of CTA 0 / thread 0 (profiles/probes/lstm_timeline.py)
#define TC_MARK(i) do { if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[i] = gtimer_ns(); } while (0)
#define TC_MARK_S(i) do { if (p.dbg && blockIdx.x == 0 && tid == 0 && s == 6) p.dbg[i] = gtimer_ns(); } while (0)

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
// DVAE_LSTM_A_TMEM: keep the resident weights in TENSOR memory instead of shared memory (tcgen05.mma with the A operand
// in TMEM): an MMA then reads only its 512-byte state tile from SMEM instead of 4 KB of weights, which is what paces
// the N = 16 MMAs of the SMEM variant (profiles/probes/umma_small_n.cu).  TMEM columns: [0, 64) accumulators,
// [64, 192) A_hi (k pair j of gate row m at lane m, column 64 + j), [192, 320) A_lo.
#ifndef DVAE_LSTM_A_TMEM
#define DVAE_LSTM_A_TMEM 1
#endif
// shared memory map (bytes): resident weights, then one block per row group
constexpr int kF_A = 0;                          // A_hi [128 x 256 fp16] 64 KB, A_lo 64 KB
constexpr int kF_APlane = 65536;
constexpr int kF_GRP = DVAE_LSTM_A_TMEM ? 0 : 131072;   // per-group blocks start here (no SMEM weights in the TMEM variant)
constexpr int kF_BPlane = kNB * kH * 2;          // one plane of the state operand: NB x 256 fp16 = 8 KB
constexpr int kG_B = 0;                          // [2 buffers][2 planes][NB x 256 fp16]
constexpr int kG_SG = 4 * kF_BPlane;             // activated gates [4][NB][32] fp32
constexpr int kG_LEN = kG_SG + 4 * kNB * 32 * 4; // int [NB]
constexpr int kG_BAR = kG_LEN + kNB * 4;         // full[2], mma_done
constexpr int kG_BYTES = kG_BAR + 64;
// per-group block for NB batch rows per group (the constants above are the NB = 16 instance; the kernels re-derive them)
__host__ __device__ constexpr int fwd_group_bytes(int NB) { return 4 * (NB * kH * 2) + 4 * NB * 32 * 4 + NB * 4 + 64; }
constexpr int fwd_smem_bytes(int G, int NB) { return kF_GRP + G * fwd_group_bytes(NB) + 16; }
// operand layouts (no swizzle, K-major; a "chunk" is 8 fp16 = 16 bytes, a core matrix is 8 rows x 1 chunk = 128 B)
//   A: chunk(m, kc) at (m >> 3) * 4096 + kc * 128 + (m & 7) * 16      SBO = 4096 (next 8 rows), LBO = 128 (next chunk)
//   B: chunk(n, kc) at kc * (NB * 16) + n * 16                         SBO = 128, LBO = NB * 16
constexpr int kF_A_SBO = 4096, kF_A_LBO = 128, kB_SBO = 128, kB_LBO = kNB * 16;
constexpr int kSlice = 4 * kB_LBO;               // this CTA's 4 K-chunks (32 units) of one plane: 1 KB
constexpr int kT_AHI = 64, kT_ALO = 192, kT_FWD_COLS = DVAE_LSTM_A_TMEM ? 512 : 64;

// G row groups of NB = 16 batch rows share the resident weights; each group is 16 warps with its own operand
// buffers, accumulators and barriers and runs the step loop independently, so one group's state exchange and gate
// math overlap the other group's MMAs on the (single) tensor pipe.
// NB = batch rows per group: 16, or 32 (one group of 32 warps) -- an N = 32 MMA takes the tensor pipe as long as an N = 16 one,
// so a bidirectional layer at B = 128 (8 clusters of 32 rows) runs ONE group per cluster instead of two that share the MMA
// issue slot, the barriers and the exchange
template <int G, int NB>
__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(32 * NB * G, 1) lstm_tc_fwd_kernel(PersistFwdArgs p) {
  constexpr int kNB = NB, kGT = 32 * NB;
  constexpr int kF_BPlane = kNB * kH * 2, kG_B = 0, kG_SG = 4 * kF_BPlane, kG_LEN = kG_SG + 4 * kNB * 32 * 4, kG_BAR = kG_LEN + kNB * 4,
                kG_BYTES = kG_BAR + 64, kB_LBO = kNB * 16, kSlice = 4 * kB_LBO;
  static_assert(kG_BYTES == fwd_group_bytes(NB), "shared-memory layout");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, grp = tid / kGT, gtid = tid % kGT, warp = gtid >> 5, lane = tid & 31, B = p.B, T = p.T;
  const uint32_t sbase = smem_u32(smem), gbase = sbase + kF_GRP + grp * kG_BYTES;
  uint8_t* gsm = smem + kF_GRP + grp * kG_BYTES;
  float* sg = reinterpret_cast<float*>(gsm + kG_SG);
  int* slen = reinterpret_cast<int*>(gsm + kG_LEN);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gsm + kG_BAR);        // [0],[1] = full, [2] = mma_done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kF_GRP + G * kG_BYTES);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kCS;
  const int d = p.d_off + cid / p.n_slices, slice = cid % p.n_slices;
  const int b0 = (slice * G + grp) * kNB, u0 = rank * kUPC;
  TC_MARK(0);

  if (gtid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (tid < 32) tmem_alloc(tmem_slot, DVAE_LSTM_A_TMEM ? kT_FWD_COLS : 32 * G);
  // resident weights: gate row m = g*32 + u  <-  W_hh[g*H + u0 + u][:], split into fp16 hi / lo planes
  if (!DVAE_LSTM_A_TMEM) {
    const float* W = p.w_hh[d];
    for (int it = tid; it < 128 * 32; it += kGT * G) {
      const int m = it >> 5, kc = it & 31;
      const int g = m >> 5, u = m & 31;
      const float4* src = reinterpret_cast<const float4*>(W + (int64_t)(g * kH + u0 + u) * kH + kc * 8);
      const float4 x0 = src[0], x1 = src[1];
      const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      __half hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_f16(xs[e], hi[e], lo[e]);
      const uint32_t off = (uint32_t)(m >> 3) * kF_A_SBO + kc * kF_A_LBO + (m & 7) * 16;
      sts128(sbase + kF_A + off, make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7])));
      sts128(sbase + kF_A + kF_APlane + off, make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7])));
    }
  }
  if (gtid < kNB) {
    const int b = b0 + gtid;
    slen[gtid] = b < B ? (p.lengths ? (int)p.lengths[b] : T) : 0;
  }
  // roles.  phase 1 (accumulator read-out): gate q of unit u0 + lane, rows 4*cgp .. 4*cgp + 3 (TMEM lane quadrant q).
  //         phase 2 (cell update): row `warp`, unit u0 + lane -- c / h live in registers across all steps.
  const int q = warp & 3, cgp = warp >> 2, prow = warp, pb = b0 + prow;
  const int plen = pb < B ? (p.lengths ? (int)p.lengths[pb] : T) : 0;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster.sync();           // every CTA's mbarriers are initialised before any peer can signal them
  tc_fence_after();
  if (DVAE_LSTM_A_TMEM && grp == 0 && cgp < 4) {      // 16 warps: 4 TMEM lane quadrants x 4 k ranges (a 32-row group has 32)
    // weights -> tensor memory: this thread owns gate row m = 32*q + lane (its TMEM lane) and k in [64*cgp, 64*cgp + 64)
    const float* wrow = p.w_hh[d] + (int64_t)(q * kH + u0 + lane) * kH + 64 * cgp;
    const uint32_t tbase = *tmem_slot + ((uint32_t)(q * 32) << 16) + 32 * cgp;
#pragma unroll
    for (int part = 0; part < 2; ++part) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = *reinterpret_cast<const float4*>(wrow + part * 32 + 4 * j);
        __half h0, l0, h1, l1, h2, l2, h3, l3;
        split_f16(x.x, h0, l0); split_f16(x.y, h1, l1); split_f16(x.z, h2, l2); split_f16(x.w, h3, l3);
        hi[2 * j] = pack2(h0, h1); hi[2 * j + 1] = pack2(h2, h3);
        lo[2 * j] = pack2(l0, l1); lo[2 * j + 1] = pack2(l2, l3);
      }
      tmem_st16(tbase + kT_AHI + part * 16, hi);
      tmem_st16(tbase + kT_ALO + part * 16, lo);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  if (DVAE_LSTM_A_TMEM) {
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem = *tmem_slot + 2 * kNB * grp;
  const uint32_t slice_off = (uint32_t)(4 * rank) * kB_LBO;
  constexpr uint32_t kIdesc = make_idesc_f16(128, kNB);

  float c_reg = 0.f, h_reg = 0.f;
  // Write this warp's row of the new state (fp16 hi / lo, operand layout) into the CTA's own operand buffer `buf`;
  // once the group's 16 rows are in, push the CTA's slice to the 7 peers (bulk copies completing on THEIR mbarrier).
  auto publish = [&](int buf) {
    const uint32_t bhi = gbase + kG_B + (uint32_t)(buf * 2) * kF_BPlane, blo = bhi + kF_BPlane;
    __half hi, lo;
    split_f16(h_reg, hi, lo);
    const uint32_t off = slice_off + (uint32_t)(lane >> 3) * kB_LBO + prow * 16 + (lane & 7) * 2;
    sts_h(bhi + off, hi);
    sts_h(blo + off, lo);
    fence_proxy_async();         // generic-proxy stores -> visible to the bulk-copy engine and to tcgen05.mma
    tc_fence_before();
    group_sync(grp, kGT);
    if (gtid == 0) mbar_expect_tx(&bars[buf], 14 * kSlice);
    if (warp == 1 && lane < 14) {
      const int pi = lane >> 1, peer = pi + (pi >= rank ? 1 : 0), plane = lane & 1;
      const uint32_t src = (plane ? blo : bhi) + slice_off;
      bulk_copy_to_peer(mapa_u32(src, peer), src, kSlice, mapa_u32(smem_u32(&bars[buf]), peer));
    }
  };
  TC_MARK(1);
  pdl_wait();               // from here on the predecessor's outputs (gates, h0 / c0) are read
  c_reg = (p.c0 && pb < B) ? p.c0[d * p.dir0 + (int64_t)pb * p.ld0 + u0 + lane] : 0.f;
  h_reg = (p.h0 && pb < B) ? p.h0[d * p.dir0 + (int64_t)pb * p.ld0 + u0 + lane] : 0.f;
  publish(0);               // initial state = "output of step -1"

  const int gcol = q * kH + u0 + lane;
  for (int s = 0; s < T; ++s) {
    const int t = d == 0 ? s : T - 1 - s;
    const int cur = s & 1;
    TC_MARK_S(2);
    // hoisted input projection (x_t W_ih^T + b) for this thread's gate column and 4 rows
    float gx[4];
    float* grow = p.gates + (((int64_t)d * T + t) * B + b0 + 4 * cgp) * 4 * kH + gcol;
#pragma unroll
    for (int j = 0; j < 4; ++j) gx[j] = (b0 + 4 * cgp + j < B) ? grow[(int64_t)j * 4 * kH] : 0.f;
    if (warp == 0) {
      mbar_wait(&bars[cur], (s >> 1) & 1);      // all 8 slices of h_{s-1} are in operand buffer `cur`
      TC_MARK_S(3);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ahi = make_smem_desc(sbase + kF_A, kF_A_LBO, kF_A_SBO, 0);
        const uint64_t alo = make_smem_desc(sbase + kF_A + kF_APlane, kF_A_LBO, kF_A_SBO, 0);
        const uint64_t bhi = make_smem_desc(gbase + kG_B + (uint32_t)(cur * 2) * kF_BPlane, kB_LBO, kB_SBO, 0);
        const uint64_t blo = make_smem_desc(gbase + kG_B + (uint32_t)(cur * 2 + 1) * kF_BPlane, kB_LBO, kB_SBO, 0);
#pragma unroll
        for (int ks = 0; ks < kH / 16; ++ks) {
          const uint64_t da = (uint64_t)(ks * 2 * kF_A_LBO >> 4), db = (uint64_t)(ks * 2 * kB_LBO >> 4);
          if (DVAE_LSTM_A_TMEM) {
            const uint32_t ta_hi = *tmem_slot + kT_AHI + 8 * ks, ta_lo = *tmem_slot + kT_ALO + 8 * ks;
            mma_f16_ts(tmem, ta_hi, bhi + db, kIdesc, ks > 0);
            mma_f16_ts(tmem + kNB, ta_hi, blo + db, kIdesc, ks > 0);
            mma_f16_ts(tmem + kNB, ta_lo, bhi + db, kIdesc, 1);
          } else {
            mma_f16(tmem, ahi + da, bhi + db, kIdesc, ks > 0);
            mma_f16(tmem + kNB, ahi + da, blo + db, kIdesc, ks > 0);
            mma_f16(tmem + kNB, alo + da, bhi + db, kIdesc, 1);
          }
        }
        tc_commit(&bars[2]);
      }
      __syncwarp();
      TC_MARK_S(4);
      mbar_wait(&bars[2], s & 1);               // only this warp polls; the other 15 sleep in the hardware barrier
    }
    group_sync(grp, kGT);
    TC_MARK_S(5);
    tc_fence_after();
    {
      float d1[4], d2[4];
      const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + 4 * cgp;
      tmem_ld4(ta, d1);
      tmem_ld4(ta + kNB, d2);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = 4 * cgp + j;
        const float pre = d1[j] + d2[j] * kLoInv + gx[j];
        float a = (q == 2) ? tanh_fast(pre) : sigmoid_fast(pre);
        if (t >= slen[row]) a = 0.f;                  // frozen / padded row: zero gates (as the reference's packing)
        sg[(q * kNB + row) * 32 + lane] = a;
        if (b0 + row < B) grow[(int64_t)j * 4 * kH] = a;
      }
    }
    tc_fence_before();
    group_sync(grp, kGT);
    TC_MARK_S(6);
    {
      float out = 0.f;
      if (t < plen) {
        const float ig = sg[(0 * kNB + prow) * 32 + lane], fg = sg[(1 * kNB + prow) * 32 + lane];
        const float gg = sg[(2 * kNB + prow) * 32 + lane], og = sg[(3 * kNB + prow) * 32 + lane];
        c_reg = fmaf(fg, c_reg, ig * gg);
        h_reg = og * tanh_fast(c_reg);
        out = h_reg;
      }
      if (pb < B) {
        p.cs[(((int64_t)d * T + t) * B + pb) * kH + u0 + lane] = c_reg;
        p.hs[((int64_t)t * B + pb) * p.ldhs + d * kH + u0 + lane] = out;
      }
    }
    TC_MARK_S(7);
    if (s + 1 < T) publish(cur ^ 1);
    TC_MARK_S(8);
  }
  TC_MARK(9);
  if (pb < B) {
    if (p.hn) p.hn[d * p.dirn + (int64_t)pb * p.ldn + u0 + lane] = h_reg;
    if (p.cn) p.cn[d * p.dirn + (int64_t)pb * p.ldn + u0 + lane] = c_reg;
  }
  tc_fence_before();
  __syncthreads();
  cluster.sync();           // no CTA leaves while a peer's bulk copy may still read from / write to it
  if (tid < 32) tmem_dealloc(*tmem_slot, DVAE_LSTM_A_TMEM ? kT_FWD_COLS : 32 * G);
}

// ------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------
// shared memory map (bytes): one block per row group (the weights live in tensor memory)
constexpr int kB_BPlane = kNB * 128 * 2;           // one plane of the dG operand: NB x 128 fp16 = 4 KB
constexpr int kB_Tile = kNB * 32 * 4;              // one [NB][32] fp32 partial-dh tile: 2 KB
constexpr int kBG_B = 0;                           // dG operand [2 planes]
constexpr int kBG_STAGE = kBG_B + 2 * kB_BPlane;   // outgoing partial dh [2 bufs][8 peers][NB][32] fp32
constexpr int kBG_RED = kBG_STAGE + 2 * 8 * kB_Tile;  // incoming partials [2 bufs][8 source CTAs][NB][32] fp32
constexpr int kBG_INV = kBG_RED + 2 * 8 * kB_Tile;    // float inv_scale[NB]
constexpr int kBG_BAR = kBG_INV + kNB * 4;            // red_full[2], mma_done
constexpr int kBG_BYTES = kBG_BAR + 64;
__host__ __device__ constexpr int bwd_group_bytes(int NB) { return 2 * (NB * 128 * 2) + 2 * (2 * 8 * NB * 32 * 4) + NB * 4 + 64; }
constexpr int bwd_smem_bytes(int G, int NB) { return G * bwd_group_bytes(NB) + 16; }
// tensor memory columns: [0, 128) accumulators (row group g: 64 g + 32 half + {0: D1, 16: D2}),
// [128, 256) A_hi (half h at 128 + 64 h: k pair j of unit 128 h + lane at column j), [256, 384) A_lo
constexpr int kTB_AHI = 128, kTB_ALO = 256, kTB_COLS = 512;

template <int G, int NB>
__global__ void __cluster_dims__(kCS, 1, 1) __launch_bounds__(32 * NB * G, 1) lstm_tc_bwd_kernel(PersistBwdArgs p) {
  constexpr int kNB = NB, kGT = 32 * NB, kB_LBO = kNB * 16;
  constexpr int kB_BPlane = kNB * 128 * 2, kB_Tile = kNB * 32 * 4, kBG_B = 0, kBG_STAGE = kBG_B + 2 * kB_BPlane,
                kBG_RED = kBG_STAGE + 2 * 8 * kB_Tile, kBG_INV = kBG_RED + 2 * 8 * kB_Tile, kBG_BAR = kBG_INV + kNB * 4,
                kBG_BYTES = kBG_BAR + 64;
  static_assert(kBG_BYTES == bwd_group_bytes(NB), "shared-memory layout");
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, grp = tid / kGT, gtid = tid % kGT, warp = gtid >> 5, lane = tid & 31, B = p.B, T = p.T;
  uint8_t* gsm = smem + grp * kBG_BYTES;
  const uint32_t gbase = smem_u32(gsm);
  float* stage = reinterpret_cast<float*>(gsm + kBG_STAGE);
  float* red = reinterpret_cast<float*>(gsm + kBG_RED);
  float* inv_s = reinterpret_cast<float*>(gsm + kBG_INV);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gsm + kBG_BAR);        // [0],[1] = red_full, [2] = mma_done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + G * kBG_BYTES);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kCS;
  const int d = p.d_off + cid / p.n_slices, slice = cid % p.n_slices;
  const int b0 = (slice * G + grp) * kNB, u0 = rank * kUPC;

  if (gtid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_barrier_init();
  }
  if (tid < 32) tmem_alloc(tmem_slot, kTB_COLS);
  if (gtid < kNB) inv_s[gtid] = 1.f;
  // roles.  read-out: TMEM lane quadrant q (units 32q + lane of each 128-unit half), rows 4*cgp .. 4*cgp + 3.
  //         cell backward: row `warp`, unit u0 + lane -- carried dh / dc live in registers.
  const int q = warp & 3, cgp = warp >> 2, prow = warp, pb = b0 + prow;
  const int plen = pb < B ? (p.lengths ? (int)p.lengths[pb] : T) : 0;
  tc_fence_before();
  __syncthreads();
  cluster.sync();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;
  if (grp == 0 && cgp < 4) {      // 16 warps: 4 TMEM lane quadrants x the 4 gates
    // resident weights, transposed, into tensor memory: A[m = unit j][k = g*32 + u] = W_hh[g*H + u0 + u][j].
    // This thread owns TMEM lane 32q + lane of both 128-unit halves and the k range of gate cgp (32 values).
    const float* W = p.w_hh[d] + (int64_t)(cgp * kH + u0) * kH;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j = 128 * half + 32 * q + lane;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        __half h0, l0, h1, l1;
        split_f16(W[(int64_t)(2 * e) * kH + j], h0, l0);
        split_f16(W[(int64_t)(2 * e + 1) * kH + j], h1, l1);
        hi[e] = pack2(h0, h1);
        lo[e] = pack2(l0, l1);
      }
      const uint32_t tb = tmem0 + ((uint32_t)(q * 32) << 16) + 64 * half + 16 * cgp;
      tmem_st16(tb + kTB_AHI, hi);
      tmem_st16(tb + kTB_ALO, lo);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem0 + 4 * kNB * grp;
  constexpr uint32_t kIdesc = make_idesc_f16(128, kNB);
  const int nsteps = T + ((p.d_h0 || p.d_c0) ? 1 : 0);
  unsigned amax_run = 0;           // bit pattern of max |dG| over this warp's row, all steps
  pdl_wait();                      // from here on the predecessor's outputs (d_hs, d_hn / d_cn) are read
  if (p.zero_buf) {                // d_x of this layer: cleared here, spread over the grid, instead of by a memset node
    float4* z4 = reinterpret_cast<float4*>(p.zero_buf);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < p.zero_n4; i += (int64_t)gridDim.x * blockDim.x)
      z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float carry = (p.d_hn && pb < B) ? p.d_hn[d * p.dirn + (int64_t)pb * p.ldn + u0 + lane] : 0.f;
  float dc = (p.d_cn && pb < B) ? p.d_cn[d * p.dirn + (int64_t)pb * p.ldn + u0 + lane] : 0.f;

  for (int s = 0; s < nsteps; ++s) {
    const bool final_ = (s == T);
    const int t = d == 0 ? T - 1 - s : s;
    const int t_prev = d == 0 ? t - 1 : t + 1;
    const int buf = s & 1;
    // prefetch the cell-backward operands of this thread's (row, unit)
    const bool live = !final_ && t < plen;
    float gi[4] = {0.f, 0.f, 0.f, 0.f}, cc = 0.f, cpv = 0.f, dhs = 0.f;
    float* gp = p.gates + (((int64_t)d * T + (final_ ? 0 : t)) * B + pb) * 4 * kH + u0 + lane;
    if (live) {
#pragma unroll
      for (int g = 0; g < 4; ++g) gi[g] = gp[g * kH];
      cc = p.cs[(((int64_t)d * T + t) * B + pb) * kH + u0 + lane];
      if (t_prev >= 0 && t_prev < T) cpv = p.cs[(((int64_t)d * T + t_prev) * B + pb) * kH + u0 + lane];
      else if (p.c0) cpv = p.c0[d * p.dir0 + (int64_t)pb * p.ld0 + u0 + lane];
      if (p.d_hs) dhs = p.d_hs[((int64_t)t * B + pb) * p.lddhs + d * kH + u0 + lane];
    }
    float rec = 0.f;
    if (s > 0) {
      // partial dh[rows, all 256 units] = dG_{t_next}[rows, own 128 gate rows] * W_hh[own gate rows, :]
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          const uint64_t bhi = make_smem_desc(gbase + kBG_B, kB_LBO, kB_SBO, 0);
          const uint64_t blo = make_smem_desc(gbase + kBG_B + kB_BPlane, kB_LBO, kB_SBO, 0);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t d1 = tmem + half * 2 * kNB, d2 = d1 + kNB;
#pragma unroll
            for (int ks = 0; ks < 128 / 16; ++ks) {
              const uint64_t db = (uint64_t)(ks * 2 * kB_LBO >> 4);
              const uint32_t ahi = tmem0 + kTB_AHI + 64 * half + 8 * ks, alo = tmem0 + kTB_ALO + 64 * half + 8 * ks;
              mma_f16_ts(d1, ahi, bhi + db, kIdesc, ks > 0);
              mma_f16_ts(d2, ahi, blo + db, kIdesc, ks > 0);
              mma_f16_ts(d2, alo, bhi + db, kIdesc, 1);
            }
          }
          tc_commit(&bars[2]);
        }
        __syncwarp();
        mbar_wait(&bars[2], (s - 1) & 1);
      }
      group_sync(grp, kGT);
      tc_fence_after();
      // read-out: each partial value goes to the staging tile of the CTA that owns its unit
      float* st_out = stage + (size_t)buf * 8 * kNB * 32;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float d1[4], d2[4];
        const uint32_t ta = tmem + ((uint32_t)(q * 32) << 16) + half * 2 * kNB + 4 * cgp;
        tmem_ld4(ta, d1);
        tmem_ld4(ta + kNB, d2);
        float* dst = st_out + (size_t)(half * 4 + q) * kNB * 32 + (4 * cgp) * 32 + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j * 32] = (d1[j] + d2[j] * kLoInv) * inv_s[4 * cgp + j];
      }
      fence_proxy_async();
      tc_fence_before();
      group_sync(grp, kGT);
      if (gtid == 0) mbar_expect_tx(&bars[buf], 7 * kB_Tile);
      if (warp == 1 && lane < 7) {
        const int peer = lane + (lane >= rank ? 1 : 0);
        const uint32_t src = gbase + kBG_STAGE + (uint32_t)(buf * 8 + peer) * kB_Tile;
        const uint32_t dst = gbase + kBG_RED + (uint32_t)(buf * 8 + rank) * kB_Tile;
        bulk_copy_to_peer(mapa_u32(dst, peer), src, kB_Tile, mapa_u32(smem_u32(&bars[buf]), peer));
      }
      if (warp == 0) mbar_wait(&bars[buf], ((s - 1) >> 1) & 1);
      group_sync(grp, kGT);
      const float* rin = red + (size_t)buf * 8 * kNB * 32 + prow * 32 + lane;
      rec = st_out[(size_t)rank * kNB * 32 + prow * 32 + lane];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k != rank) rec += rin[(size_t)k * kNB * 32];
    }
    if (final_) {
      if (pb < B) {
        if (p.d_h0) p.d_h0[d * p.dird0 + (int64_t)pb * p.ldd0 + u0 + lane] = carry + rec;
        if (p.d_c0) p.d_c0[d * p.dird0 + (int64_t)pb * p.ldd0 + u0 + lane] = dc;
      }
      break;
    }
    // LSTM cell backward -> gate gradients; then the next step's operand (fp16 hi / lo, per-row power-of-two scale)
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (live) {
      const float ig = gi[0], fg = gi[1], gg = gi[2], og = gi[3];
      const float dh = carry + rec + dhs;
      const float tcv = tanh_fast(cc);
      const float dct = fmaf(dh * og, 1.f - tcv * tcv, dc);
      o[0] = dct * gg * ig * (1.f - ig);
      o[1] = dct * cpv * fg * (1.f - fg);
      o[2] = dct * ig * (1.f - gg * gg);
      o[3] = dh * tcv * og * (1.f - og);
      carry = 0.f;
      dc = dct * fg;
    } else {
      carry += rec;              // frozen row: pass the carried gradient through
    }
    if (pb < B) {
#pragma unroll
      for (int g = 0; g < 4; ++g) gp[g * kH] = o[g];
    }
    // row scale 2^(13 - floor(log2(max |dG[row, own 128 gate rows]|))): exact, undone when the accumulator is read
    const float am = fmaxf(fmaxf(fabsf(o[0]), fabsf(o[1])), fmaxf(fabsf(o[2]), fabsf(o[3])));
    const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(am));
    amax_run = max(amax_run, mx);
    int se = 267 - (int)(mx >> 23);
    se = se < 1 ? 1 : (se > 253 ? 253 : se);
    const float sc = __uint_as_float((unsigned)se << 23);
    if (lane == 0) inv_s[prow] = __uint_as_float((unsigned)(254 - se) << 23);
    const uint32_t bhi = gbase + kBG_B, blo = bhi + kB_BPlane;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      __half hi, lo;
      split_f16(o[g] * sc, hi, lo);
      const uint32_t off = (uint32_t)(g * 4 + (lane >> 3)) * kB_LBO + prow * 16 + (lane & 7) * 2;
      sts_h(bhi + off, hi);
      sts_h(blo + off, lo);
    }
    fence_proxy_async();
    tc_fence_before();
    group_sync(grp, kGT);
  }
  // max |dG| of this CTA -> its own slot (plain store: nothing to initialise, no memset node in front of the kernel);
  // the GEMMs that use dG as an operand take the maximum over the grid's slots
  __shared__ unsigned s_amax;
  if (tid == 0) s_amax = 0u;
  __syncthreads();
  if (lane == 0 && amax_run) atomicMax(&s_amax, amax_run);
  tc_fence_before();
  __syncthreads();
  if (p.amax_out && tid == 0) p.amax_out[blockIdx.x] = s_amax;
  cluster.sync();
  if (tid < 32) tmem_dealloc(tmem0, kTB_COLS);
}

}  // namespace

bool tc_lstm_supported(int H) {
  const char* e = getenv("DVAE_LSTM_IMPL");
  return H == kH && !(e && !strcmp(e, "simt"));
}

// launch with programmatic stream serialization (DVAE_PDL=0: plain launch)
template <class Args>
static cudaError_t launch_pdl(void (*kernel)(Args), int grid, int block, size_t smem, cudaStream_t st, const Args& args) {
  const char* e = getenv("DVAE_PDL");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (e && e[0] == '0') ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, args);
}

template <int G, int NB>
static int launch_tc_fwd(const PersistFwdArgs& a, cudaStream_t st) {
  static bool ready = false;
  if (!ready) {
    DVAE_CUDA(cudaFuncSetAttribute(lstm_tc_fwd_kernel<G, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem_bytes(G, NB)));
    ready = true;
  }
  PersistFwdArgs b = a;
  b.n_slices = ceil_div(a.B, NB * G);
  b.d_off = 0;
  DVAE_CUDA(launch_pdl(lstm_tc_fwd_kernel<G, NB>, kCS * b.n_slices * a.D, 32 * NB * G, fwd_smem_bytes(G, NB), st, b));
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

// rows per cluster: 16 (one 16-row group) while D * ceil(B / 16) clusters fit the 15 a B200 keeps resident; else 32, as two
// 16-row groups whose phases interleave (one group's exchange and gate math under the other's MMAs).  DVAE_LSTM_GROUPS=32
// selects ONE 32-row group instead (N = 32 MMAs, half the barriers): measured SLOWER at cfg 2 -- 3.54 vs 3.26 us per step of
// the bidirectional encoder, step 1.415 vs 1.293 ms -- because a single group serialises its phases; kept as a tested A/B.
static int tc_layout(int B, int D) {
  const char* e = getenv("DVAE_LSTM_GROUPS");
  const int want = e ? atoi(e) : 0;
  if (want == 1 || want == 2 || want == 32) return want;
  return D * ceil_div(B, kNB) <= 15 ? 1 : 2;
}

int tc_lstm_fwd(const PersistFwdArgs& a, cudaStream_t st) {
  // at most 15 clusters of 8 are co-resident on a B200 (profiles/probes/cluster_occupancy.cu).  One row group per
  // cluster (lowest step latency) while everything fits in one wave; two row groups per cluster otherwise, which
  // also keeps both directions of a bidirectional layer in ONE launch at B = 128.
  const int lay = tc_layout(a.B, a.D);
  return lay == 32 ? launch_tc_fwd<1, 32>(a, st) : (lay == 2 ? launch_tc_fwd<2, 16>(a, st) : launch_tc_fwd<1, 16>(a, st));
}

template <int G, int NB>
static int launch_tc_bwd(const PersistBwdArgs& a, cudaStream_t st) {
  static bool ready = false;
  if (!ready) {
    DVAE_CUDA(cudaFuncSetAttribute(lstm_tc_bwd_kernel<G, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem_bytes(G, NB)));
    ready = true;
  }
  PersistBwdArgs b = a;
  b.n_slices = ceil_div(a.B, NB * G);
  b.d_off = 0;
  DVAE_CUDA(launch_pdl(lstm_tc_bwd_kernel<G, NB>, kCS * b.n_slices * a.D, 32 * NB * G, bwd_smem_bytes(G, NB), st, b));
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

int tc_lstm_bwd(const PersistBwdArgs& a, cudaStream_t st) {
  const int lay = tc_layout(a.B, a.D);
  return lay == 32 ? launch_tc_bwd<1, 32>(a, st) : (lay == 2 ? launch_tc_bwd<2, 16>(a, st) : launch_tc_bwd<1, 16>(a, st));
}

// CTAs of the backward launch = entries written to PersistBwdArgs::amax_out
int tc_lstm_bwd_ctas(int B, int D) { return kCS * ceil_div(B, tc_layout(B, D) == 1 ? kNB : 2 * kNB) * D; }

}  // namespace dvae
