// Tensor-core GEMM, second generation: C[M,N] = A . B^T with fp32 operands in HBM, computed by tcgen05.mma.kind::f16
// on fp16 (hi, lo) operand planes produced on the fly:
//     hi = fp16(x * s)            lo = fp16((x * s - hi) * 2^11)            (s = optional power-of-two operand scale)
//     D1 += A_hi * B_hi           D2 += A_hi * B_lo + A_lo * B_hi           C = (D1 + D2 * 2^-11) / (s_a * s_b)
// 22 mantissa bits per operand and fp32 accumulation in TMEM: fp32-grade results (forward-loss parity 1e-5, identical
// arg-max) for THREE K=16 MMAs per 16 k, where the 3xTF32 kernel (tc_gemm.cu) needs six K=8 ones and twice the
// shared-memory bytes per k.  Shared-memory bandwidth is what bounds both kernels, so the work is organised to
// touch SMEM as little as possible:
//   warp 5    (TMA): cp.async.bulk.tensor (no swizzle, zero fill handles every tail) of raw fp32 A / B k-blocks into a
//              3-stage landing ring, so 96 KB of loads are in flight per SM (bytes in flight, not SMEM bandwidth or
//              MMA issue, is what bounds these fp32-operand GEMMs: profiles/probes/tc16_timeline.py).
//   warps 6-13 (converters): conflict-free LDS of the landed tile, scale + split, 16-byte stores into the UMMA
//              no-swizzle K-major core-matrix layout (padded so that the stores are conflict-free too).  MN-major
//              sources (transposed operands of the backward GEMMs) are transposed on the way, so the MMA only ever
//              sees K-major tiles.
//   warp 4    (MMA): elect.sync'ed single-thread issue from a warp-uniform branch (descriptors stay in uniform
//              registers), 5-stage ring, tcgen05.commit frees ring slots and publishes accumulators.
//   warps 0-3 (epilogue): double-buffered accumulators (2 x (D1, D2) = 512 TMEM columns) so read-out, bias /
//              activation / softmax statistics and the global stores of tile i overlap the main loop of tile i+1.
// Modes are those of tc_gemm.cu: 0 plain linear (bias, tanh, beta, split-K), 1 vocab-CE forward partials (online
// log-softmax / arg-max / Gumbel-max sampling), 2 softmax-gradient chunk.
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include "tc_gemm.cuh"
#include "tc_gemm16.cuh"

namespace dvae {
namespace tc16 {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3 /* operand-plane ring */, RAW_STAGES = 2 /* TMA landing ring */;
constexpr int T_LBO = 160, T_SBO = 4 * T_LBO;      // chunk(row, kc) at (row >> 3) * 640 + kc * 160 + (row & 7) * 16:
                                                   // core matrices (8 rows x 16 B) 160 B apart along K so that the 16
                                                   // chunks a warp stores per instruction spread over all banks
constexpr int PLANE = (BM / 8) * T_SBO;            // one fp16 plane of a 128 x 32 tile: 10 KB
constexpr int STAGE_BYTES = 4 * PLANE;             // A_hi, A_lo, B_hi, B_lo
constexpr int RAW_TILE = BM * BK * 4;              // one landed fp32 tile: 16 KB
constexpr int RAW_BYTES = 2 * RAW_TILE;            // A, B
constexpr int EPI_WARPS = 8, MMA_WARP = 4, TMA_WARP = 5, PROD_WARP0 = 6, PROD_WARPS = 16, EPI2_WARP0 = PROD_WARP0 + PROD_WARPS;
// warps 0-3: epilogue of tile columns 0-63; warps 22-25 (22 % 4 == 2: TMEM lane quarter = warp & 3): columns 64-127
constexpr int NUM_THREADS = (EPI2_WARP0 + 4) * 32;               // 832: <= 78 registers per thread
constexpr int EPI_SCRATCH_BYTES = 8 * 32 * 33 * 4;
// [operand rings][epilogue transpose scratch][bias tiles][barriers]: the scratch directly follows the operand region so
// that the A-stationary ring can grow into it when the epilogue does not use it (mode 1)
constexpr int OFF_RAW = STAGES * STAGE_BYTES, OFF_OPER_END = OFF_RAW + RAW_STAGES * RAW_BYTES, OFF_SCRATCH = OFF_OPER_END;
// Pre-split mode (operands already split into fp16 planes in global memory, blocked by core matrix): no landing ring and
// no converters -- TMA delivers UMMA-ready dense tiles (core matrices 128 B apart) straight into a 5-stage operand ring.
constexpr int PS_PLANE = BM * BK * 2, PS_STAGE_BYTES = 4 * PS_PLANE, PS_STAGES = 5, MAX_STAGES = 5;
constexpr int PS_LBO = 128, PS_SBO = 512;
// A-stationary variant (K <= 256, no split-K): the CTA's A rows (both planes, all k-blocks: 128 KB) are loaded once and
// stay in shared memory for the CTA's whole run of N tiles; only B tiles stream through a 3-stage ring.  With 128 x 128
// tiles and K = 256 the operand loads (256 KB per tile) otherwise make the kernel L2-bandwidth bound.
constexpr int AS_MAX_KB = 8, AS_A_BYTES = AS_MAX_KB * 2 * PS_PLANE, AS_STAGE_BYTES = 2 * PS_PLANE;
constexpr int AS_STAGES = 3, AS_STAGES_NOSCRATCH = 5;      // mode 1 has no transpose scratch: the ring grows into it
static_assert(AS_A_BYTES + AS_STAGES * AS_STAGE_BYTES <= OFF_OPER_END, "A-stationary layout must fit in the operand region");
static_assert(AS_STAGES_NOSCRATCH <= MAX_STAGES, "barrier arrays");
static_assert(PS_STAGES * PS_STAGE_BYTES <= OFF_OPER_END, "pre-split ring must fit in the plane + landing rings");
constexpr int OFF_BIAS = OFF_SCRATCH + EPI_SCRATCH_BYTES;          // per epilogue warp: the 64 bias values of its tile columns
constexpr int OFF_BAR = OFF_BIAS + 8 * 64 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
static_assert(AS_A_BYTES + AS_STAGES_NOSCRATCH * AS_STAGE_BYTES <= OFF_BIAS, "A-stationary mode-1 ring must end before the bias tiles");
constexpr int TMEM_COLS = 512;
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

__device__ __forceinline__ unsigned long long gtime16() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DBG16(i) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[i] = gtime16(); } while (0)

__device__ __forceinline__ float scale_from_amax(const uint32_t* amax_bits, float static_scale) {
  if (!amax_bits) return static_scale;
  // 2^(13 - floor(log2 amax)): the largest element lands in [2^13, 2^14), well inside fp16 range
  int se = 267 - (int)(*amax_bits >> 23);
  se = se < 1 ? 1 : (se > 253 ? 253 : se);
  return __uint_as_float((unsigned)se << 23);
}

// (x0, x1) * s -> packed fp16 hi pair (returned) and lo pair; packed fp32 arithmetic (FMUL2 / FFMA2)
__device__ __forceinline__ uint32_t pack_hi_lo(float x0, float x1, float s, uint32_t& lo) {
  const float2 x = __fmul2_rn(make_float2(x0, x1), make_float2(s, s));
  const __half2 h = __float22half2_rn(x);
  const float2 hf = __half22float2(h);
  // (x - hi) * 2^11 = x * 2^11 - hi * 2^11 (both products exact)
  const float2 r = __ffma2_rn(x, make_float2(kLoScale, kLoScale), __fmul2_rn(hf, make_float2(-kLoScale, -kLoScale)));
  const __half2 l = __float22half2_rn(r);
  lo = *reinterpret_cast<const uint32_t*>(&l);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One landed k-block (128 rows x 32 k, fp32) of one operand -> 8 registers per converter thread (512 threads).
//   K-major source: the tile is [128 rows][32 k]; piece p = ptid + 512*i (i < 2) is the float4 at row p >> 3, k 4*(p & 7):
//     a warp reads 512 contiguous bytes per instruction (conflict-free).
//   MN-major source: the tile is [32 k][128 rows]; thread = (row ptid & 127, k chunk ptid >> 7), 8 scalar reads with the
//     warp's lanes on consecutive rows (conflict-free): the transposition happens in registers.
__device__ __forceinline__ void load_tile(uint32_t raw, int mn_major, int ptid, float (&v)[8]) {
  if (!mn_major) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint4 t = lds128(raw + (uint32_t)(ptid + 512 * i) * 16);
      v[4 * i] = __uint_as_float(t.x); v[4 * i + 1] = __uint_as_float(t.y);
      v[4 * i + 2] = __uint_as_float(t.z); v[4 * i + 3] = __uint_as_float(t.w);
    }
  } else {
    const uint32_t src = raw + (uint32_t)(8 * (ptid >> 7)) * (BM * 4) + (uint32_t)(ptid & 127) * 4;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = lds32(src + j * (BM * 4));
  }
}

// scale, split into fp16 (hi, lo) and store into the stage's operand planes (hi plane at `plane_hi`, lo at + PLANE)
__device__ __forceinline__ void store_tile(uint32_t plane_hi, int mn_major, int ptid, float s, const float (&v)[8]) {
  if (!mn_major) {
    const int odd = ptid & 1;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int pc = ptid + 512 * i, row = pc >> 3, kc = (pc & 7) >> 1;
      uint32_t lo0, lo1;
      const uint32_t hi0 = pack_hi_lo(v[4 * i], v[4 * i + 1], s, lo0);
      const uint32_t hi1 = pack_hi_lo(v[4 * i + 2], v[4 * i + 3], s, lo1);
      // lanes (2j, 2j+1) hold k 0-3 / 4-7 of one 8-k chunk: the even lane assembles the hi chunk, the odd lane the lo chunk
      const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? hi0 : lo0, 1);
      const uint32_t r1 = __shfl_xor_sync(0xffffffffu, odd ? hi1 : lo1, 1);
      const uint4 chunk = odd ? make_uint4(r0, r1, lo0, lo1) : make_uint4(hi0, hi1, r0, r1);
      sts128(plane_hi + (odd ? PLANE : 0) + (uint32_t)(row >> 3) * T_SBO + (uint32_t)kc * T_LBO + (row & 7) * 16, chunk);
    }
  } else {
    const int row = ptid & 127, kc = ptid >> 7;
    const uint32_t base = plane_hi + (uint32_t)(row >> 3) * T_SBO + (uint32_t)kc * T_LBO + (row & 7) * 16;
    uint4 hi, lo;
    hi.x = pack_hi_lo(v[0], v[1], s, lo.x);
    hi.y = pack_hi_lo(v[2], v[3], s, lo.y);
    hi.z = pack_hi_lo(v[4], v[5], s, lo.z);
    hi.w = pack_hi_lo(v[6], v[7], s, lo.w);
    sts128(base, hi);
    sts128(base + PLANE, lo);
  }
}

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// X [R, K] fp32 (row stride ld) -> fp16 planes hi / lo, each blocked by UMMA core matrix:
//   plane[((r / 8) * KC + k / 8) * 64 + (r % 8) * 8 + k % 8],  KC = ceil(K / 8); rows / columns beyond R / K are zero.
__global__ void split_planes_kernel(const float* __restrict__ X, int64_t ld, int R, int K, float scale, __half* __restrict__ hi,
                                    __half* __restrict__ lo) {
  const int KC = (K + 7) / 8, RP = (R + 7) / 8 * 8;
  const int64_t total = (int64_t)RP * KC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kc = (int)(i % KC), r = (int)(i / KC);            // consecutive threads: consecutive 32-byte pieces of a row
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (r < R && kc * 8 + e < K) ? __ldg(X + (int64_t)r * ld + kc * 8 + e) : 0.f;
    uint4 h, l;
    h.x = pack_hi_lo(v[0], v[1], scale, l.x); h.y = pack_hi_lo(v[2], v[3], scale, l.y);
    h.z = pack_hi_lo(v[4], v[5], scale, l.z); h.w = pack_hi_lo(v[6], v[7], scale, l.w);
    const int64_t off = (((int64_t)(r >> 3) * KC + kc) * 64 + (r & 7) * 8);
    *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc16_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;                       // [MAX_STAGES] converters (or TMA, pre-split mode) -> MMA
  uint64_t* empty = bars + MAX_STAGES;         // [MAX_STAGES] MMA -> converters / TMA   (tcgen05.commit)
  uint64_t* raw_full = bars + 2 * MAX_STAGES;  // [RAW_STAGES] TMA -> converters (transaction bytes)
  uint64_t* raw_empty = raw_full + RAW_STAGES; // [RAW_STAGES] converters -> TMA (8 warp arrivals)
  uint64_t* tmem_full = raw_empty + RAW_STAGES;  // [2] MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;          // [2] epilogue -> MMA (4 warp arrivals)
  uint64_t* a_full = tmem_empty + 2;             // [1] A-stationary mode: all A k-blocks landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);
  float* epi_scratch = reinterpret_cast<float*>(smem + OFF_SCRATCH);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // a programmatically-launched successor (the persistent LSTM kernels) may start its prologue now; it still waits for
  // this grid to complete (griddepcontrol.wait) before it reads anything written here
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (tid == 0) DBG16(0);
  const int m0 = blockIdx.x * BM;
  const int nt0 = blockIdx.y * p.tiles_per_cta;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int nt1 = min(n_tiles, nt0 + p.tiles_per_cta);
  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int nkb = max(0, min(nkb_total, kb0 + p.kb_per_split) - kb0);

  if (tid == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&full[s], p.presplit ? 1 : PROD_WARPS);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < RAW_STAGES; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&raw_empty[s], PROD_WARPS);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_WARPS);
    }
    mbar_init(a_full, 1);
    fence_barrier_init();
  }
  if (warp == TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.presplit) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
  }
  const bool a_stat = p.presplit && nkb_total <= AS_MAX_KB && gridDim.z == 1;
  const int nstages = a_stat ? (p.mode == 1 ? AS_STAGES_NOSCRATCH : AS_STAGES) : (p.presplit ? PS_STAGES : STAGES);
  const uint32_t stage_bytes = a_stat ? AS_STAGE_BYTES : (p.presplit ? PS_STAGE_BYTES : STAGE_BYTES);
  const uint32_t plane_bytes = p.presplit ? PS_PLANE : PLANE;
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_u = smem_u32(smem);

  if (warp == TMA_WARP) {
    // ===== TMA producer: raw fp32 k-blocks into the landing ring =====
    if (lane == 0 && a_stat) {
      // A rows once (all k-blocks, both planes), then only B tiles through the ring
      mbar_expect_tx(a_full, (uint32_t)nkb * 2 * PS_PLANE);
      for (int kb = 0; kb < nkb; ++kb) {
        tma_load_3d(smem_u + kb * 2 * PS_PLANE, &tmA, 0, kb * (BK / 8), m0 / 8, a_full);
        tma_load_3d(smem_u + kb * 2 * PS_PLANE + PS_PLANE, &tmA2, 0, kb * (BK / 8), m0 / 8, a_full);
      }
      int stage = 0, phase = 0;
      for (int nt = nt0; nt < nt1; ++nt) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t dst = smem_u + AS_A_BYTES + stage * AS_STAGE_BYTES;
          mbar_expect_tx(&full[stage], AS_STAGE_BYTES);
          const int rb = (nt * BN + p.b_row0) / 8;
          tma_load_3d(dst, &tmB, 0, kb * (BK / 8), rb, &full[stage]);
          tma_load_3d(dst + PS_PLANE, &tmB2, 0, kb * (BK / 8), rb, &full[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (lane == 0 && p.presplit) {
      // operand planes straight into the operand ring: 4 x 8 KB dense tiles per k-block
      int stage = 0, phase = 0;
      for (int nt = nt0; nt < nt1; ++nt) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          const uint32_t dst = smem_u + stage * PS_STAGE_BYTES;
          mbar_expect_tx(&full[stage], PS_STAGE_BYTES);
          const int kc0 = (kb0 + kb) * (BK / 8), ra = m0 / 8, rb = (nt * BN + p.b_row0) / 8;
          tma_load_3d(dst, &tmA, 0, kc0, ra, &full[stage]);
          tma_load_3d(dst + PS_PLANE, &tmA2, 0, kc0, ra, &full[stage]);
          tma_load_3d(dst + 2 * PS_PLANE, &tmB, 0, kc0, rb, &full[stage]);
          tma_load_3d(dst + 3 * PS_PLANE, &tmB2, 0, kc0, rb, &full[stage]);
          if (++stage == PS_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (lane == 0) {
      int rs = 0, rphase = 0;
      for (int nt = nt0; nt < nt1; ++nt) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&raw_empty[rs], rphase ^ 1);
          uint8_t* dst = smem + OFF_RAW + rs * RAW_BYTES;
          mbar_expect_tx(&raw_full[rs], RAW_BYTES);
          const int k0 = (kb0 + kb) * BK;
          if (!p.a_mn) tma_load_2d(dst, &tmA, k0, m0, &raw_full[rs]);
          else tma_load_2d(dst, &tmA, m0, k0, &raw_full[rs]);
          if (!p.b_mn) tma_load_2d(dst + RAW_TILE, &tmB, k0, nt * BN, &raw_full[rs]);
          else tma_load_2d(dst + RAW_TILE, &tmB, nt * BN, k0, &raw_full[rs]);
          if (++rs == RAW_STAGES) { rs = 0; rphase ^= 1; }
        }
      }
    }
  } else if (warp >= PROD_WARP0 && warp < EPI2_WARP0) {
    // ===== converters: landed fp32 tile -> (scale, split) -> fp16 operand planes =====
    const int ptid = tid - PROD_WARP0 * 32;
    const float sa = scale_from_amax(p.a_amax, p.a_scale), sb = scale_from_amax(p.b_amax, p.b_scale);
    const int n_items = p.presplit ? 0 : (nt1 - nt0) * nkb;
    int stage = 0, phase = 0, rs = 0, rphase = 0;
    for (int it = 0; it < n_items; ++it) {
      const bool mark = ptid == 0 && it == 12;
      if (mark) DBG16(2);
      mbar_wait(&raw_full[rs], rphase);
      if (mark) DBG16(3);
      float va[8], vb[8];
      const uint32_t raw = smem_u + OFF_RAW + rs * RAW_BYTES;
      load_tile(raw, p.a_mn, ptid, va);
      load_tile(raw + RAW_TILE, p.b_mn, ptid, vb);
      mbar_wait(&empty[stage], phase ^ 1);
      const uint32_t st = smem_u + stage * STAGE_BYTES;
      store_tile(st, p.a_mn, ptid, sa, va);
      store_tile(st + 2 * PLANE, p.b_mn, ptid, sb, vb);
      if (mark) DBG16(4);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full[stage]);
        mbar_arrive(&raw_empty[rs]);
      }
      if (mark) DBG16(5);
      if (ptid == 0 && it == 13) DBG16(6);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
      if (++rs == RAW_STAGES) { rs = 0; rphase ^= 1; }
    }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // f16 x f16 -> f32
    int stage = 0, phase = 0, tile = 0;
    if (a_stat && nkb > 0) mbar_wait(a_full, 0);
    for (int nt = nt0; nt < nt1 && nkb > 0; ++nt, ++tile) {
      const int acc = tile & 1;
      if (tile >= 2) {
        mbar_wait(&tmem_empty[acc], ((tile >> 1) - 1) & 1);
        tc_fence_after();
      }
      const uint32_t d1 = tmem_base + acc * 256, d2 = d1 + 128;
      for (int kb = 0; kb < nkb; ++kb) {
        if (lane == 0 && tile == 0 && kb == 12) DBG16(8);
        if (lane == 0 && tile == 1 && kb < 4) DBG16(20 + 2 * kb);
        mbar_wait(&full[stage], phase);
        if (lane == 0 && tile == 1 && kb < 4) DBG16(21 + 2 * kb);
        if (lane == 0 && tile == 0 && kb == 12) DBG16(9);
        if (lane == 0 && tile == 0 && kb == 13) DBG16(11);
        // no tcgen05.fence here: the operands arrive through the async proxy (TMA) or behind the converters'
        // fence.proxy.async, and the mbarrier wait orders them; a tcgen05.fence::after_thread_sync per k-block makes the
        // issuing thread wait for the MMAs already in flight (0.45 us instead of 0.2 us per k-block)
        if (elect_one()) {
          const uint32_t lbo = p.presplit ? PS_LBO : T_LBO, sbo = p.presplit ? PS_SBO : T_SBO;
          // operand bases: ring stage [A_hi, A_lo, B_hi, B_lo], or stationary A (k-block kb) + ring stage [B_hi, B_lo]
          const uint32_t sa = a_stat ? smem_u + kb * 2 * PS_PLANE : smem_u + stage * stage_bytes;
          const uint32_t sb = a_stat ? smem_u + AS_A_BYTES + stage * AS_STAGE_BYTES : sa + 2 * plane_bytes;
          const uint64_t ahi = make_smem_desc(sa, lbo, sbo, 0), alo = make_smem_desc(sa + plane_bytes, lbo, sbo, 0);
          const uint64_t bhi = make_smem_desc(sb, lbo, sbo, 0), blo = make_smem_desc(sb + plane_bytes, lbo, sbo, 0);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 2 * lbo >> 4);
            const uint32_t accum = (kb | k) ? 1u : 0u;
            if (p.dbg_skip_epilogue == 2) continue;        // probes only: operand delivery in isolation
            mma_f16(d1, ahi + adv, bhi + adv, idesc, accum);
            mma_f16(d2, ahi + adv, blo + adv, idesc, accum);
            mma_f16(d2, alo + adv, bhi + adv, idesc, 1u);
          }
          tc_commit(&empty[stage]);
          if (kb == nkb - 1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (lane == 0 && tile == 0 && kb == 12) DBG16(10);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue: 8 warps = 4 TMEM lane quarters x 2 column halves of the tile =====
    const int quarter = warp & 3, ehalf = warp >= EPI2_WARP0 ? 1 : 0, ew = ehalf * 4 + quarter;
    const int row = m0 + quarter * 32 + lane;      // output row owned by this thread (TMEM lane)
    const bool row_ok = row < p.M;
    const float oscale = p.alpha * (p.alpha_dev ? *p.alpha_dev : 1.f) /
                         (scale_from_amax(p.a_amax, p.a_scale) * scale_from_amax(p.b_amax, p.b_scale));
    // per-row state of the fused vocabulary epilogues
    float rm = -INFINITY, rs = 0.f, rt = 0.f, rav = -INFINITY, row_lse = 0.f, row_scale = 0.f;
    int rai = 0x7fffffff, tgt = -1;
    if (p.mode != 0 && row_ok) {
      const int b = row % p.B, tpos = row / p.B + 1;
      if (p.targets) tgt = (int)p.targets[(int64_t)b * p.tgt_stride_b + tpos];
      if (p.mode == 2) {
        row_lse = p.lse[row];
        row_scale = (tpos < p.lengths[b]) ? (p.grad_scale ? p.grad_scale[0] : 1.f) / (float)p.B : 0.f;
        tgt -= p.v0;
      }
    }
    int tile = 0;
    for (int nt = nt0; nt < nt1 && nkb > 0; ++nt, ++tile) {
      const int n0 = nt * BN, acc = tile & 1;
      mbar_wait(&tmem_full[acc], (tile >> 1) & 1);
      tc_fence_after();
      if (tid == 0 && tile < 4) DBG16(14 + 2 * tile);
      // this warp's 64 bias values (vocabulary modes) -> shared memory: broadcast reads instead of 64 global loads
      float* sbias = reinterpret_cast<float*>(smem + OFF_BIAS) + ew * 64;
      if (p.mode != 0) {
        __syncwarp();
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int col = n0 + ehalf * 64 + h2 * 32 + lane;
          sbias[h2 * 32 + lane] = col < p.N ? __ldg(p.bias + col) : 0.f;
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int c = 2 * ehalf; c < 2 * ehalf + 2; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N || p.dbg_skip_epilogue) continue;                       // warp-uniform (probes: 1 = no epilogue, 2 = no MMAs either)
        const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256 + c * 32;
        if (p.mode != 1) {
          // modes 0 / 2 store a [32 rows x 32 cols] chunk: transpose it through padded shared memory so each
          // store instruction covers 32 consecutive columns of one row (coalesced) instead of 32 different rows
          const uint32_t sc = smem_u32(epi_scratch) + ew * (32 * 33 * 4);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float v[16], w[16];
            tmem_ld16(ta + hh * 16, v);
            tmem_ld16(ta + 128 + hh * 16, w);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = fmaf(w[j], kLoInv, v[j]) * oscale;
              if (p.mode == 2) {
                const int col = col0 + hh * 16 + j;
                x = (row_scale == 0.f || col >= p.N) ? 0.f
                    : (__expf(x + sbias[(c & 1) * 32 + hh * 16 + j] - row_lse) - (col == tgt ? 1.f : 0.f)) * row_scale;
              }
              sts32(sc + (lane * 33 + hh * 16 + j) * 4, x);
            }
          }
          __syncwarp();
          const int col = col0 + lane;
          const int r0 = m0 + quarter * 32;
          const int nr = min(32, p.M - r0);              // valid rows of this warp's 32-row band
          if (col < p.N && nr > 0) {
            const bool split = gridDim.z > 1;
            float badd = 0.f;
            if (p.mode == 0 && (!split || blockIdx.z == 0)) {
              if (p.bias) badd += __ldg(p.bias + col);
              if (p.bias2) badd += __ldg(p.bias2 + col);
            }
            float* cp = p.C + (int64_t)r0 * p.ldc + col;
            const int64_t ldc = p.ldc;
            const uint32_t src = sc + lane * 4;
            if (p.mode == 2 || (!split && p.act == 0 && p.beta == 0.f)) {
              if (nr == 32) {
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) cp[rr * ldc] = lds32(src + rr * 132) + badd;
              } else {
                for (int rr = 0; rr < nr; ++rr) cp[rr * ldc] = lds32(src + rr * 132) + badd;
              }
            } else if (split) {
              for (int rr = 0; rr < nr; ++rr) atomicAdd(cp + rr * ldc, lds32(src + rr * 132) + badd);   // C pre-scaled by beta
            } else {
              const float beta = p.beta;
              const bool do_tanh = p.act == 1;
              if (nr == 32 && beta != 0.f) {
                // read-modify-write of C: all 32 loads in flight before the first dependent store
#pragma unroll 1
                for (int r8 = 0; r8 < 32; r8 += 8) {
                  float old[8];
#pragma unroll
                  for (int rr = 0; rr < 8; ++rr) old[rr] = cp[(r8 + rr) * ldc];
#pragma unroll
                  for (int rr = 0; rr < 8; ++rr) {
                    float x = lds32(src + (r8 + rr) * 132) + badd;
                    if (do_tanh) x = tanhf(x);
                    cp[(r8 + rr) * ldc] = fmaf(beta, old[rr], x);
                  }
                }
              } else {
                for (int rr = 0; rr < nr; ++rr) {
                  float x = lds32(src + rr * 132) + badd;
                  if (do_tanh) x = tanhf(x);
                  if (beta != 0.f) x = fmaf(beta, cp[rr * ldc], x);
                  cp[rr * ldc] = x;
                }
              }
            }
          }
          __syncwarp();
          continue;
        }
        // mode 1: online log-softmax statistics of this row over the tile's columns (logits never leave registers)
        const bool sample = p.gumbel_seed != nullptr;
        const uint64_t gseed = sample ? *p.gumbel_seed : 0;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v[16], w[16];
          tmem_ld16(ta + hh * 16, v);
          tmem_ld16(ta + 128 + hh * 16, w);
          if (!row_ok) continue;
          // 16 logits of this row: everything below is a tree or independent per element (the serial running-max /
          // arg-max / sum chains of the obvious loop made this epilogue, not the MMAs, the kernel's critical path)
          const int base = col0 + hh * 16;
          float x[16], xs[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            x[j] = base + j < p.N ? fmaf(w[j], kLoInv, v[j]) * oscale + sbias[(c & 1) * 32 + hh * 16 + j] : -INFINITY;
          if (sample) {
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              float g[4];
              gumbel4(gseed, p.gumbel_salt, row, base + j4, (p.N + 3) >> 2, g);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) xs[j4 + jj] = x[j4 + jj] + g[jj];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) xs[j] = x[j];
          }
          float m8[8], m4[4];
#pragma unroll
          for (int j = 0; j < 8; ++j) m8[j] = fmaxf(xs[2 * j], xs[2 * j + 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) m4[j] = fmaxf(m8[2 * j], m8[2 * j + 1]);
          const float smax = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));       // max of the (noisy) scores
          if (smax > rav) {                                  // columns ascend: ties keep the earlier index
            unsigned eq = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) eq |= (xs[j] == smax ? 1u : 0u) << j;
            rav = smax;
            rai = base + __ffs(eq) - 1;
          }
          if (tgt >= base && tgt < base + 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (base + j == tgt) rt = x[j];
          }
          float tmax = smax;
          if (sample) {                                      // the softmax statistics use the noise-free logits
#pragma unroll
            for (int j = 0; j < 8; ++j) m8[j] = fmaxf(x[2 * j], x[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 4; ++j) m4[j] = fmaxf(m8[2 * j], m8[2 * j + 1]);
            tmax = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          }
          if (tmax > rm) { rs *= __expf(rm - tmax); rm = tmax; }
          float e8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) e8[j] = __expf(x[2 * j] - rm) + __expf(x[2 * j + 1] - rm);
          rs += ((e8[0] + e8[1]) + (e8[2] + e8[3])) + ((e8[4] + e8[5]) + (e8[6] + e8[7]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (tid == 0 && tile < 4) DBG16(15 + 2 * tile);
    }
    if (p.mode == 1 && row_ok) {      // the two column halves are separate vocabulary splits for the finalize kernel
      const int64_t sp = (int64_t)blockIdx.y * 2 + ehalf;
      *reinterpret_cast<float4*>(p.part + (sp * p.M + row) * 4) = make_float4(rm, rs, rt, rav);
      p.part_idx[sp * p.M + row] = rai;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
  if (tid == 0) DBG16(1);
}

// ---- host ------------------------------------------------------------------------------------------
bool supported(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, int M, int N, int K) {
  return enabled() && shape_ok(A, lda, trans_a, B, ldb, trans_b, M, N, K);
}

bool enabled() {
  // default; DVAE_GEMM_IMPL=tf32 keeps the 3xTF32 kernel (tc_gemm.cu) for A/B runs, =simt the fp32 SIMT kernels
  const char* e = getenv("DVAE_GEMM_IMPL");
  return !(e && (e[0] == 't' || e[0] == 's'));
}

bool shape_ok(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, int M, int N, int K) {
  // TMA: 16-byte aligned bases and row strides
  if (((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) || lda % 4 || ldb % 4) return false;
  return M >= 1 && N >= 1 && K >= 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  return fn;
}
// un-swizzled 2-D fp32 map over a row-major [outer, inner] matrix with row stride ld: K-major operands land as
// [128 rows][32 k] (box 32 x 128), MN-major ones as [32 k][128 rows] (box 128 x 32)
static int make_map(CUtensorMap* m, const float* base, int64_t ld, int mn_major, int rows, int K) {
  EncodeTiledFn fn = encode_fn();
  DVAE_REQUIRE(fn != nullptr, "tc16_gemm: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)(mn_major ? rows : K), (cuuint64_t)(mn_major ? K : rows)};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)(mn_major ? BM : BK), (cuuint32_t)(mn_major ? BK : BM)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DVAE_REQUIRE(r == CUDA_SUCCESS, "tc16_gemm: cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%lld", (int)r, rows, K, (long long)ld);
  return DVAE_OK;
}

// 3-D map over one fp16 plane in the blocked layout of split_planes: dim0 = one core matrix (64 elements, 128 B),
// dim1 = k chunks, dim2 = 8-row groups; a box of 64 x 4 x 16 is a dense UMMA-ready 128 x 32 operand tile
static int make_plane_map(CUtensorMap* m, const void* plane, int rows, int K) {
  EncodeTiledFn fn = encode_fn();
  DVAE_REQUIRE(fn != nullptr, "tc16_gemm: cuTensorMapEncodeTiled is not available from the driver");
  const int KC = ceil_div(K, 8), RG = ceil_div(rows, 8);
  cuuint64_t dims[3] = {64, (cuuint64_t)KC, (cuuint64_t)RG};
  cuuint64_t strides[2] = {128, (cuuint64_t)KC * 128};
  cuuint32_t box[3] = {64, BK / 8, BM / 8};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(plane), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DVAE_REQUIRE(r == CUDA_SUCCESS, "tc16_gemm: plane tensor map failed (%d) rows=%d K=%d", (int)r, rows, K);
  return DVAE_OK;
}

int64_t plane_floats(int R, int K) { return (int64_t)ceil_div(R, 8) * 8 * ceil_div(K, 8) * 8; }

bool presplit_enabled() {
  const char* e = getenv("DVAE_VOCAB_PRESPLIT");
  return !(e && e[0] == '0');
}

int split_planes(const float* X, int64_t ld, int R, int K, float scale, void* planes, cudaStream_t st) {
  DVAE_REQUIRE(X && planes && R > 0 && K > 0, "tc16 split_planes: bad argument");
  const int64_t n = plane_floats(R, K);                 // fp16 elements per plane = floats of both planes / ... (2 B each)
  __half* hi = reinterpret_cast<__half*>(planes);
  __half* lo = hi + n;
  const int64_t work = (int64_t)ceil_div(R, 8) * 8 * ceil_div(K, 8);
  int blocks = ceil_div(work, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  split_planes_kernel<<<blocks, 256, 0, st>>>(X, ld, R, K, scale, hi, lo);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

static int launch(const Params& p, dim3 grid, cudaStream_t st) {
  static bool ready = false;
  if (!ready) {
    DVAE_CUDA(cudaFuncSetAttribute(tc16_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ready = true;
  }
  CUtensorMap ma, mb, ma2, mb2;
  memset(&ma2, 0, sizeof(ma2));
  memset(&mb2, 0, sizeof(mb2));
  int rc;
  if (p.presplit) {
    const __half* ah = reinterpret_cast<const __half*>(p.a_planes);
    const __half* bh = reinterpret_cast<const __half*>(p.b_planes);
    if ((rc = make_plane_map(&ma, ah, p.a_rows, p.K))) return rc;
    if ((rc = make_plane_map(&ma2, ah + plane_floats(p.a_rows, p.K), p.a_rows, p.K))) return rc;
    if ((rc = make_plane_map(&mb, bh, p.b_rows, p.K))) return rc;
    if ((rc = make_plane_map(&mb2, bh + plane_floats(p.b_rows, p.K), p.b_rows, p.K))) return rc;
  } else {
    if ((rc = make_map(&ma, p.A, p.lda, p.a_mn, p.M, p.K))) return rc;
    if ((rc = make_map(&mb, p.Bm, p.ldb, p.b_mn, p.N, p.K))) return rc;
  }
  tc16_gemm_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, ma2, mb2, p);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

__global__ void tc16_scale_rows_kernel(float* C, int64_t ldc, int M, int N, float beta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  float* c = C + (i / N) * ldc + (i % N);
  *c = beta == 0.f ? 0.f : *c * beta;
}

static void apply_hints(Params& p, const GemmHints& h) {
  if (const char* e = getenv("DVAE_TC_DBG")) p.dbg = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
  if (const char* e = getenv("DVAE_TC_SKIP_EPILOGUE")) p.dbg_skip_epilogue = atoi(e);     // probes only: main-loop speed in isolation
  p.a_amax = h.a_amax_bits; p.b_amax = h.b_amax_bits;
  p.a_scale = h.a_scale; p.b_scale = h.b_scale;
  p.alpha = 1.f; p.alpha_dev = nullptr;
}

int linear(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C, int64_t ldc, int M,
           int N, int K, const float* bias, const float* bias2, float beta, int act, const GemmHints& hints, cudaStream_t st) {
  Params p = {};
  p.A = A; p.lda = lda; p.Bm = B; p.ldb = ldb;
  p.M = M; p.N = N; p.K = K; p.a_mn = trans_a ? 1 : 0; p.b_mn = trans_b ? 1 : 0; p.tiles_per_cta = 1;
  p.C = C; p.ldc = ldc; p.bias = bias; p.bias2 = bias2; p.beta = beta; p.act = act; p.mode = 0;
  apply_hints(p, hints);
  // split-K when the output has too few tiles to occupy the 148 SMs and K is deep (weight-gradient shapes)
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN), nkb = ceil_div(K, BK);
  int splits = 1;
  if (act == 0 && tiles * 2 <= 148 && nkb >= 16) {
    splits = 148 / tiles;
    if (splits > nkb / 8) splits = nkb / 8;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = ceil_div(nkb, splits);
  splits = ceil_div(nkb, p.kb_per_split);
  if (splits > 1 && beta != 1.f) {
    if (beta == 0.f && ldc == N) {
      DVAE_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    } else {
      tc16_scale_rows_kernel<<<ceil_div((int64_t)M * N, 256), 256, 0, st>>>(C, ldc, M, N, beta);
      DVAE_LAUNCH_CHECK();
    }
  }
  // more tiles than SMs: two N tiles per CTA, so the second tile's main loop hides the first one's epilogue
  if (splits == 1 && tiles > 148 && ceil_div(N, BN) >= 2) p.tiles_per_cta = 2;
  return launch(p, dim3(ceil_div(M, BM), ceil_div(ceil_div(N, BN), p.tiles_per_cta), splits), st);
}

int ce_partials(const float* h, int64_t ldh, int N, int B, int H, int V, const float* w, const float* bias,
                const int64_t* targets, int64_t tgt_stride_b, const int64_t* lengths, int tiles_per_split, int nsplit,
                float* part, int* part_idx, const uint64_t* gumbel_seed, uint32_t gumbel_salt, const void* h_planes,
                const void* w_planes, cudaStream_t st) {
  Params p = {};
  if (h_planes && w_planes) { p.presplit = 1; p.a_planes = h_planes; p.b_planes = w_planes; p.a_rows = N; p.b_rows = V; }
  p.A = h; p.lda = ldh; p.Bm = w; p.ldb = H;
  p.M = N; p.N = V; p.K = H; p.tiles_per_cta = tiles_per_split; p.bias = bias; p.mode = 1;
  p.kb_per_split = ceil_div(H, BK);
  p.targets = targets; p.tgt_stride_b = tgt_stride_b; p.lengths = lengths; p.B = B; p.part = part; p.part_idx = part_idx;
  p.gumbel_seed = gumbel_seed; p.gumbel_salt = gumbel_salt;
  apply_hints(p, GemmHints());
  return launch(p, dim3(ceil_div(N, BM), nsplit, 1), st);
}

int softmax_grad(const float* h, int64_t ldh, int N, int B, int H, int V, int v0, int vc, const float* w, const float* bias,
                 const int64_t* targets, int64_t tgt_stride_b, const int64_t* lengths, const float* lse,
                 const float* grad_scale, float* P, int64_t ldp, const void* h_planes, const void* w_planes, cudaStream_t st) {
  Params p = {};
  if (h_planes && w_planes) { p.presplit = 1; p.a_planes = h_planes; p.b_planes = w_planes; p.a_rows = N; p.b_rows = V; p.b_row0 = v0; }
  p.A = h; p.lda = ldh; p.Bm = w + (int64_t)v0 * H; p.ldb = H;
  // a run of vocabulary tiles per CTA so the epilogue of one tile overlaps the main loop of the next; ~one wave of CTAs
  const int row_tiles = ceil_div(N, BM), col_tiles = ceil_div(vc, BN);
  int per_cta = ceil_div(col_tiles, max(1, 148 / row_tiles));
  if (per_cta < 1) per_cta = 1;
  p.M = N; p.N = vc; p.K = H; p.tiles_per_cta = per_cta; p.bias = bias + v0; p.mode = 2;
  p.kb_per_split = ceil_div(H, BK);
  p.targets = targets; p.tgt_stride_b = tgt_stride_b; p.lengths = lengths; p.B = B; p.lse = lse; p.grad_scale = grad_scale;
  p.v0 = v0; p.C = P; p.ldc = ldp;
  apply_hints(p, GemmHints());
  return launch(p, dim3(row_tiles, ceil_div(col_tiles, per_cta), 1), st);
}

}  // namespace tc16
}  // namespace dvae

extern "C" int dvae_tc16_linear(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C,
                                int64_t ldc, int M, int N, int K, const float* bias, const float* bias2, float beta, int act,
                                float a_scale, float b_scale, const uint32_t* a_amax_bits, const uint32_t* b_amax_bits,
                                void* stream) {
  DVAE_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "dvae_tc16_linear: bad argument");
  DVAE_REQUIRE(act == 0 || act == 1, "dvae_tc16_linear: unknown activation %d", act);
  DVAE_REQUIRE(dvae::tc16::shape_ok(A, lda, trans_a, B, ldb, trans_b, M, N, K),
               "dvae_tc16_linear: operands must be 16-byte aligned with ld %% 4 == 0 (TMA)");
  dvae::GemmHints h;
  h.a_scale = a_scale; h.b_scale = b_scale; h.a_amax_bits = a_amax_bits; h.b_amax_bits = b_amax_bits;
  return dvae::tc16::linear(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, bias2, beta, act, h, (cudaStream_t)stream);
}
