// Tensor-core GEMM, second generation: C[M,N] = A . B^T with fp32 operands in HBM, computed by tcgen05.mma.kind::f16
// on fp16 (hi, lo) operand planes produced on the fly:
//     hi = fp16(x * s)            lo = fp16((x * s - hi) * 2^11)            (s = optional power-of-two operand scale)
//     D1 += A_hi * B_hi           D2 += A_hi * B_lo + A_lo * B_hi           C = (D1 + D2 * 2^-11) / (s_a * s_b)
// 22 mantissa bits per operand and fp32 accumulation in TMEM: fp32-grade results (forward-loss parity 1e-5, identical
// arg-max) for THREE K=16 MMAs per 16 k, where the 3xTF32 kernel (tc_gemm.cu) needs six K=8 ones and twice the
// shared-memory bytes per k.  Shared-memory bandwidth is what bounds both kernels, so the work is organised to
// touch SMEM as little as possible:
//   warps 5-12 (producers): coalesced / sector-exact global loads straight into registers (no TMA landing zone),
//              scale + split, 16-byte stores into the UMMA no-swizzle K-major core-matrix layout (bank-conflict free).
//              MN-major sources (transposed operands of the backward GEMMs) are transposed in registers on the way,
//              so the MMA only ever sees K-major tiles.  Per 128x128x32 k-block: 32 KB written, 48 KB read by the MMAs
//              (the 3xTF32 kernel moves 224 KB).
//   warp 4    (MMA): elect.sync'ed single-thread issue from a warp-uniform branch (descriptors stay in uniform
//              registers), 5-stage ring, tcgen05.commit frees ring slots and publishes accumulators.
//   warps 0-3 (epilogue): double-buffered accumulators (2 x (D1, D2) = 512 TMEM columns) so read-out, bias /
//              activation / softmax statistics and the global stores of tile i overlap the main loop of tile i+1.
// Modes are those of tc_gemm.cu: 0 plain linear (bias, tanh, beta, split-K), 1 vocab-CE forward partials (online
// log-softmax / arg-max / Gumbel-max sampling), 2 softmax-gradient chunk.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_gemm.cuh"
#include "tc_gemm16.cuh"

namespace dvae {
namespace tc16 {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 5;
constexpr int T_LBO = 160, T_SBO = 4 * T_LBO;      // chunk(row, kc) at (row >> 3) * 640 + kc * 160 + (row & 7) * 16:
                                                   // core matrices (8 rows x 16 B) 160 B apart along K so that the 16
                                                   // chunks a warp stores per instruction spread over all banks
constexpr int PLANE = (BM / 8) * T_SBO;            // one fp16 plane of a 128 x 32 tile: 10 KB
constexpr int STAGE_BYTES = 4 * PLANE;             // A_hi, A_lo, B_hi, B_lo
constexpr int EPI_WARPS = 4, MMA_WARP = 4, PROD_WARP0 = 5, PROD_WARPS = 8;
constexpr int NUM_THREADS = (PROD_WARP0 + PROD_WARPS) * 32;      // 416
constexpr int EPI_SCRATCH_BYTES = 4 * 32 * 33 * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + EPI_SCRATCH_BYTES + 1024;
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

__device__ __forceinline__ uint32_t pack_hi_lo(float x0, float x1, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((x0 - hf.x) * kLoScale, (x1 - hf.y) * kLoScale);
  lo = *reinterpret_cast<const uint32_t*>(&l);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One k-block (128 rows x 32 k) of one operand -> 16 registers per producer thread (256 threads).
//   K-major source ([rows, K] row-major): piece p = ptid + 256*i (i < 4) is the float4 at row p >> 3, k 4*(p & 7): the 8
//     lanes of a row read 128 contiguous bytes (4 fully used lines per warp instruction).
//   MN-major source ([K, rows] row-major): thread = (row ptid & 127, k half ptid >> 7), 16 scalar loads, each
//     coalesced across the warp (lanes hold consecutive rows): the transposition happens in registers.
__device__ __forceinline__ void load_tile(const float* __restrict__ X, int64_t ld, int mn_major, int ptid, int row0, int rows,
                                          int k0, int K, float (&v)[16]) {
  if (!mn_major) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pc = ptid + 256 * i, row = row0 + (pc >> 3), k = k0 + 4 * (pc & 7);
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows && k < K) t = __ldcg(reinterpret_cast<const float4*>(X + (int64_t)row * ld + k));   // K % 4 == 0
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
    const int row = row0 + (ptid & 127), kk = k0 + 16 * (ptid >> 7);
    const float* src = X + (int64_t)kk * ld + row;
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (row < rows && kk + j < K) ? __ldcg(src + (int64_t)j * ld) : 0.f;
  }
}

// scale, split into fp16 (hi, lo) and store into the stage's operand planes (hi plane at `plane_hi`, lo at + PLANE)
__device__ __forceinline__ void store_tile(uint32_t plane_hi, int mn_major, int ptid, float s, const float (&v)[16]) {
  if (!mn_major) {
    const int odd = ptid & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pc = ptid + 256 * i, row = pc >> 3, kc = (pc & 7) >> 1;
      uint32_t lo0, lo1;
      const uint32_t hi0 = pack_hi_lo(v[4 * i] * s, v[4 * i + 1] * s, lo0);
      const uint32_t hi1 = pack_hi_lo(v[4 * i + 2] * s, v[4 * i + 3] * s, lo1);
      // lanes (2j, 2j+1) hold k 0-3 / 4-7 of one 8-k chunk: the even lane assembles the hi chunk, the odd lane the lo chunk
      const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? hi0 : lo0, 1);
      const uint32_t r1 = __shfl_xor_sync(0xffffffffu, odd ? hi1 : lo1, 1);
      const uint4 chunk = odd ? make_uint4(r0, r1, lo0, lo1) : make_uint4(hi0, hi1, r0, r1);
      sts128(plane_hi + (odd ? PLANE : 0) + (uint32_t)(row >> 3) * T_SBO + (uint32_t)kc * T_LBO + (row & 7) * 16, chunk);
    }
  } else {
    const int row = ptid & 127, half = ptid >> 7;
    const uint32_t base = plane_hi + (uint32_t)(row >> 3) * T_SBO + (uint32_t)(2 * half) * T_LBO + (row & 7) * 16;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint4 hi, lo;
      hi.x = pack_hi_lo(v[8 * c + 0] * s, v[8 * c + 1] * s, lo.x);
      hi.y = pack_hi_lo(v[8 * c + 2] * s, v[8 * c + 3] * s, lo.y);
      hi.z = pack_hi_lo(v[8 * c + 4] * s, v[8 * c + 5] * s, lo.z);
      hi.w = pack_hi_lo(v[8 * c + 6] * s, v[8 * c + 7] * s, lo.w);
      sts128(base + c * T_LBO, hi);
      sts128(base + c * T_LBO + PLANE, lo);
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) tc16_gemm_kernel(Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES] producers -> MMA   (8 warp arrivals)
  uint64_t* empty = bars + STAGES;             // [STAGES] MMA -> producers   (tcgen05.commit)
  uint64_t* tmem_full = bars + 2 * STAGES;     // [2] MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2] epilogue -> MMA (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* epi_scratch = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) DBG16(0);
  const int m0 = blockIdx.x * BM;
  const int nt0 = blockIdx.y * p.tiles_per_cta;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int nt1 = min(n_tiles, nt0 + p.tiles_per_cta);
  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int nkb = max(0, min(nkb_total, kb0 + p.kb_per_split) - kb0);

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], PROD_WARPS);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_u = smem_u32(smem);

  if (warp >= PROD_WARP0) {
    // ===== producers: global -> registers -> (scale, split) -> operand planes =====
    const int ptid = tid - PROD_WARP0 * 32;
    const float sa = scale_from_amax(p.a_amax, p.a_scale), sb = scale_from_amax(p.b_amax, p.b_scale);
    const int n_items = (nt1 - nt0) * nkb;
    // Three rotating register buffers: the loads of k-blocks i+1 and i+2 are in flight while k-block i is converted
    // and stored (one L2 round trip per k-block would otherwise bound the kernel).
    float a0[16], b0[16], a1[16], b1[16], a2[16], b2[16];
    auto fetch = [&](int it, float (&va)[16], float (&vb)[16]) {
      if (it < n_items) {
        const int nt = nt0 + it / nkb, kb = it % nkb;
        load_tile(p.A, p.lda, p.a_mn, ptid, m0, p.M, (kb0 + kb) * BK, p.K, va);
        load_tile(p.Bm, p.ldb, p.b_mn, ptid, nt * BN, p.N, (kb0 + kb) * BK, p.K, vb);
      }
    };
    int stage = 0, phase = 0, done = 0;
    auto commit = [&](const float (&va)[16], const float (&vb)[16]) {
      const bool mark = ptid == 0 && done == 12;
      if (mark) DBG16(2);
      mbar_wait(&empty[stage], phase ^ 1);
      if (mark) DBG16(3);
      const uint32_t st = smem_u + stage * STAGE_BYTES;
      store_tile(st, p.a_mn, ptid, sa, va);
      store_tile(st + 2 * PLANE, p.b_mn, ptid, sb, vb);
      if (mark) DBG16(4);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[stage]);
      if (mark) DBG16(5);
      if (ptid == 0 && done == 13) DBG16(6);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
      ++done;
    };
    fetch(0, a0, b0);
    fetch(1, a1, b1);
    for (int it = 0; it < n_items; it += 3) {
      fetch(it + 2, a2, b2);
      commit(a0, b0);
      if (it + 1 < n_items) {
        fetch(it + 3, a0, b0);
        commit(a1, b1);
      }
      if (it + 2 < n_items) {
        fetch(it + 4, a1, b1);
        commit(a2, b2);
      }
    }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // f16 x f16 -> f32
    int stage = 0, phase = 0, tile = 0;
    for (int nt = nt0; nt < nt1 && nkb > 0; ++nt, ++tile) {
      const int acc = tile & 1;
      if (tile >= 2) {
        mbar_wait(&tmem_empty[acc], ((tile >> 1) - 1) & 1);
        tc_fence_after();
      }
      const uint32_t d1 = tmem_base + acc * 256, d2 = d1 + 128;
      for (int kb = 0; kb < nkb; ++kb) {
        if (lane == 0 && tile == 0 && kb == 12) DBG16(8);
        mbar_wait(&full[stage], phase);
        if (lane == 0 && tile == 0 && kb == 12) DBG16(9);
        if (lane == 0 && tile == 0 && kb == 13) DBG16(11);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u + stage * STAGE_BYTES;
          const uint64_t ahi = make_smem_desc(sa, T_LBO, T_SBO, 0), alo = make_smem_desc(sa + PLANE, T_LBO, T_SBO, 0);
          const uint64_t bhi = make_smem_desc(sa + 2 * PLANE, T_LBO, T_SBO, 0), blo = make_smem_desc(sa + 3 * PLANE, T_LBO, T_SBO, 0);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 2 * T_LBO >> 4);
            const uint32_t accum = (kb | k) ? 1u : 0u;
            mma_f16(d1, ahi + adv, bhi + adv, idesc, accum);
            mma_f16(d2, ahi + adv, blo + adv, idesc, accum);
            mma_f16(d2, alo + adv, bhi + adv, idesc, 1u);
          }
          tc_commit(&empty[stage]);
          if (kb == nkb - 1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (lane == 0 && tile == 0 && kb == 12) DBG16(10);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===== epilogue (warps 0..3 = TMEM lane quarters 0..3) =====
    const int quarter = warp;
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
      if (tid == 0 && tile == 0) DBG16(14);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        {
          float w[32];
          const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256 + c * 32;
          tmem_ld32(ta, v);
          tmem_ld32(ta + 128, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(w[j], kLoInv, v[j]) * oscale;
        }
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) continue;                       // warp-uniform
        if (p.mode != 1) {
          // modes 0 / 2 store a [32 rows x 32 cols] chunk: transpose it through padded shared memory so each
          // store instruction covers 32 consecutive columns of one row (coalesced) instead of 32 different rows
          const uint32_t sc = smem_u32(epi_scratch) + warp * (32 * 33 * 4);
          if (p.mode == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              v[j] = (row_scale == 0.f || col >= p.N) ? 0.f
                     : (expf(v[j] + __ldg(p.bias + col) - row_lse) - (col == tgt ? 1.f : 0.f)) * row_scale;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) sts32(sc + (lane * 33 + j) * 4, v[j]);
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
              for (int rr = 0; rr < nr; ++rr) {
                float x = lds32(src + rr * 132) + badd;
                if (do_tanh) x = tanhf(x);
                if (beta != 0.f) x = fmaf(beta, cp[rr * ldc], x);
                cp[rr * ldc] = x;
              }
            }
          }
          __syncwarp();
          continue;
        }
        if (!row_ok) continue;
        {
          // online log-softmax statistics of this row over the tile's columns (logits never leave registers)
          float tmax = -INFINITY;
          const bool sample = p.gumbel_seed != nullptr;
          const uint64_t gseed = sample ? *p.gumbel_seed : 0;
#pragma unroll
          for (int j4 = 0; j4 < 32; j4 += 4) {
            float g[4] = {0.f, 0.f, 0.f, 0.f};
            if (sample) gumbel4(gseed, p.gumbel_salt, row, col0 + j4, (p.N + 3) >> 2, g);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = j4 + jj, col = col0 + j;
              float x = col < p.N ? v[j] + __ldg(p.bias + col) : -INFINITY;
              v[j] = x;
              tmax = fmaxf(tmax, x);
              const float xs = x + g[jj];
              if (xs > rav) { rav = xs; rai = col; }       // columns ascend, so ties keep the first index
              if (col == tgt) rt = x;
            }
          }
          if (tmax > rm) { rs *= expf(rm - tmax); rm = tmax; }
#pragma unroll
          for (int j = 0; j < 32; ++j) rs += expf(v[j] - rm);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (tid == 0 && tile == 0) DBG16(15);
    }
    if (p.mode == 1 && row_ok) {
      *reinterpret_cast<float4*>(p.part + ((int64_t)blockIdx.y * p.M + row) * 4) = make_float4(rm, rs, rt, rav);
      p.part_idx[(int64_t)blockIdx.y * p.M + row] = rai;
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
  // Opt-in (DVAE_GEMM_IMPL=f16): on the cfg-2 shapes this kernel and the 3xTF32 one (tc_gemm.cu) both sit at about
  // 1 us per 32 KB k-block -- bound by fp32-operand bytes in flight per SM, not by SMEM traffic or MMA issue
  // (profiles/probes/tc16_perf.py, tc16_timeline.py) -- and the TMA ring of the older kernel hides slightly more latency.
  const char* e = getenv("DVAE_GEMM_IMPL");
  return e && e[0] == 'f';
}

bool shape_ok(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, int M, int N, int K) {
  // K-major operands are read with 16-byte loads along K
  if (!trans_a && ((reinterpret_cast<uintptr_t>(A) & 15) || lda % 4 || K % 4)) return false;
  if (!trans_b && ((reinterpret_cast<uintptr_t>(B) & 15) || ldb % 4 || K % 4)) return false;
  return M >= 1 && N >= 1 && K >= 1;
}

static int launch(const Params& p, dim3 grid, cudaStream_t st) {
  static bool ready = false;
  if (!ready) {
    DVAE_CUDA(cudaFuncSetAttribute(tc16_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ready = true;
  }
  tc16_gemm_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
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
  return launch(p, dim3(ceil_div(M, BM), ceil_div(N, BN), splits), st);
}

int ce_partials(const float* h, int64_t ldh, int N, int B, int H, int V, const float* w, const float* bias,
                const int64_t* targets, int64_t tgt_stride_b, const int64_t* lengths, int tiles_per_split, int nsplit,
                float* part, int* part_idx, const uint64_t* gumbel_seed, uint32_t gumbel_salt, cudaStream_t st) {
  Params p = {};
  p.A = h; p.lda = ldh; p.Bm = w; p.ldb = H;
  p.M = N; p.N = V; p.K = H; p.tiles_per_cta = tiles_per_split; p.bias = bias; p.mode = 1;
  p.kb_per_split = ceil_div(H, BK);
  p.targets = targets; p.tgt_stride_b = tgt_stride_b; p.lengths = lengths; p.B = B; p.part = part; p.part_idx = part_idx;
  p.gumbel_seed = gumbel_seed; p.gumbel_salt = gumbel_salt;
  apply_hints(p, GemmHints());
  return launch(p, dim3(ceil_div(N, BM), nsplit, 1), st);
}

int softmax_grad(const float* h, int64_t ldh, int N, int B, int H, int v0, int vc, const float* w, const float* bias,
                 const int64_t* targets, int64_t tgt_stride_b, const int64_t* lengths, const float* lse,
                 const float* grad_scale, float* P, int64_t ldp, cudaStream_t st) {
  Params p = {};
  p.A = h; p.lda = ldh; p.Bm = w + (int64_t)v0 * H; p.ldb = H;
  p.M = N; p.N = vc; p.K = H; p.tiles_per_cta = 1; p.bias = bias + v0; p.mode = 2;
  p.kb_per_split = ceil_div(H, BK);
  p.targets = targets; p.tgt_stride_b = tgt_stride_b; p.lengths = lengths; p.B = B; p.lse = lse; p.grad_scale = grad_scale;
  p.v0 = v0; p.C = P; p.ldc = ldp;
  apply_hints(p, GemmHints());
  return launch(p, dim3(ceil_div(N, BM), ceil_div(vc, BN), 1), st);
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
               "dvae_tc16_linear: K-major operands must be 16-byte aligned with ld %% 4 == 0 and K %% 4 == 0");
  dvae::GemmHints h;
  h.a_scale = a_scale; h.b_scale = b_scale; h.a_amax_bits = a_amax_bits; h.b_amax_bits = b_amax_bits;
  return dvae::tc16::linear(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, bias2, beta, act, h, (cudaStream_t)stream);
}
