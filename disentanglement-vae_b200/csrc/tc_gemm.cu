// Tensor-core GEMM for sm_100a: C[M,N] = A . B^T with fp32 operands in HBM, computed on the 5th-gen
// tensor cores as 3xTF32 (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) so the result keeps fp32-grade
// accuracy (forward loss parity 1e-5, identical argmax), or 1xTF32 (round-to-nearest operands).
//
// One CTA = one 128x128 output tile (or a run of tiles along N that share the A rows):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor (SWIZZLE_128B) of raw fp32 A / B k-blocks (BK = 32) into a
//              3-stage shared-memory ring, completion on mbarriers (zero fill handles M / N / K tails)
//   warps 2-5: split workers -- rewrite each landed tile in place as its tf32 "hi" part and write the
//              residual "lo" part to a twin buffer (element-wise, so the swizzle never has to be decoded),
//              fence.proxy.async, then signal the MMA warp; afterwards they are the epilogue warps
//   warp 1   : MMA issuer    -- one thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=128, K=8) from
//              shared-memory descriptors (K-major or MN-major, so transposed operands need no copy),
//              tcgen05.commit releases ring slots / publishes the accumulator
//   epilogue : tcgen05.ld 32x32b -> registers -> bias / activation / beta -> global
#include <stdlib.h>

#include "tc_gemm.cuh"

namespace dvae {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;            // 16 KB
constexpr int STAGE_BYTES = 4 * TILE_BYTES;        // A_hi, A_lo, B_hi, B_lo
constexpr int NUM_THREADS = 192;
constexpr int EPI_SCRATCH_BYTES = 4 * 32 * 33 * 4;  // per epilogue warp: one padded 32x32 fp32 chunk (transpose for coalesced stores)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */ + EPI_SCRATCH_BYTES;
constexpr int TMEM_COLS = 128;

struct Params {
  int M, N, K;
  int a_mn, b_mn;        // 1 = operand is MN-major in global memory (stored [K, rows])
  int passes;            // 3 = 3xTF32, 1 = TF32 with round-to-nearest operands
  int tiles_per_cta;     // consecutive N tiles per CTA
  float* C; int64_t ldc;
  const float* bias; const float* bias2;
  float beta; int act;
  int mode;              // 0: C = act(acc + bias) + beta*C;  1: vocab-CE forward partials;  2: softmax-gradient tile
  int kb_per_split;      // k-blocks per blockIdx.z (split-K, mode 0; partial tiles are atomically accumulated)
  // modes 1 / 2 (rows are decoder positions n = (t-1)*B + b, columns are vocabulary ids)
  const int64_t* targets; int64_t tgt_stride_b; const int64_t* lengths; int B;
  float* part; int* part_idx;            // mode 1: [nsplit][M][4] (max, sumexp, target logit, argmax value), [nsplit][M]
  const float* lse; const float* grad_scale; int v0;   // mode 2: C = P[:, v0:v0+N]
  const uint64_t* gumbel_seed; uint32_t gumbel_salt;   // mode 1: when set, the arg-max is taken over logits + Gumbel noise (sampling)
  unsigned long long* dbg;   // optional: pipeline milestone timestamps (ns) of CTA (0,0,0)
};

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DBG_MARK(i) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.dbg[i] = gtime(); } while (0)

__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];          // SWIZZLE_128B tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_raw = bars;
  uint64_t* full_split = bars + STAGES;
  uint64_t* empty = bars + 2 * STAGES;
  uint64_t* tmem_full = bars + 3 * STAGES;
  uint64_t* tmem_empty = bars + 3 * STAGES + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 2);
  float* epi_scratch = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) DBG_MARK(0);
  const int m0 = blockIdx.x * BM;
  const int nt0 = blockIdx.y * p.tiles_per_cta;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int nt1 = min(n_tiles, nt0 + p.tiles_per_cta);
  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int nkb = max(0, min(nkb_total, kb0 + p.kb_per_split) - kb0);

  if (tid == 32) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&full_split[s], 4);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) DBG_MARK(1);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int nt = nt0; nt < nt1; ++nt) {
        const int n0 = nt * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full_raw[stage], 2 * TILE_BYTES);
          if (!p.a_mn) {
            tma_load_2d(st, &tmA, (kb0 + kb) * BK, m0, &full_raw[stage]);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 32; ++j) tma_load_2d(st + j * (BK * 128), &tmA, m0 + 32 * j, (kb0 + kb) * BK, &full_raw[stage]);
          }
          if (!p.b_mn) {
            tma_load_2d(st + 2 * TILE_BYTES, &tmB, (kb0 + kb) * BK, n0, &full_raw[stage]);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 32; ++j)
              tma_load_2d(st + 2 * TILE_BYTES + j * (BK * 128), &tmB, n0 + 32 * j, (kb0 + kb) * BK, &full_raw[stage]);
          }
          if (nt == nt0 && kb == 0) DBG_MARK(2);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(BM, BN, p.a_mn, p.b_mn);
      // K-major: LBO unused (1 x 16 B), SBO = 8 rows x 128 B; advance 32 B per K=8 step inside the swizzle row.
      // MN-major (tf32 => SW128 with 32 B atoms): k-rows of 128 B, 4 k-rows per atom -> SBO = 512 B; LBO = next
      // 32-wide chunk along M/N (one TMA box = BK x 128 B); advance 8 k-rows = 1024 B per K=8 step.
      const uint32_t a_lbo = p.a_mn ? BK * 128 : 16, b_lbo = p.b_mn ? BK * 128 : 16;
      const uint32_t a_sbo = p.a_mn ? 512 : 1024, b_sbo = p.b_mn ? 512 : 1024;
      const uint32_t a_kstep = p.a_mn ? 1024 : 32, b_kstep = p.b_mn ? 1024 : 32;
      const uint32_t a_lt = p.a_mn ? 1 : 2, b_lt = p.b_mn ? 1 : 2;
      int stage = 0, phase = 0, tile = 0;
      for (int nt = nt0; nt < nt1; ++nt, ++tile) {
        if (tile > 0) {
          mbar_wait(tmem_empty, (tile - 1) & 1);
          tc_fence_after();
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_split[stage], phase);
          tc_fence_after();
          if (nt == nt0 && kb == 0) DBG_MARK(4);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES), sb = sa + 2 * TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            const uint64_t a_hi = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, a_lt);
            const uint64_t b_hi = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo, b_lt);
            if (p.passes == 3) {
              const uint64_t a_lo = make_smem_desc(sa + TILE_BYTES + k * a_kstep, a_lbo, a_sbo, a_lt);
              const uint64_t b_lo = make_smem_desc(sb + TILE_BYTES + k * b_kstep, b_lbo, b_sbo, b_lt);
              mma_tf32(tmem_base, a_lo, b_hi, idesc, (kb | k) ? 1u : 0u);
              mma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
              mma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
            } else {
              mma_tf32(tmem_base, a_hi, b_hi, idesc, (kb | k) ? 1u : 0u);
            }
          }
          tc_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tmem_full);
        if (nt == nt0) DBG_MARK(5);
      }
    }
  } else {
    // ===== split workers / epilogue (warps 2..5) =====
    const int wtid = tid - 64;
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
    const int row = m0 + quarter * 32 + lane;      // output row owned by this thread (TMEM lane)
    const bool row_ok = row < p.M;
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
    int stage = 0, phase = 0, tile = 0;
    for (int nt = nt0; nt < nt1; ++nt, ++tile) {
      const int n0 = nt * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_raw[stage], phase);
        if (wtid == 0 && nt == nt0 && kb == 0) DBG_MARK(3);
        const uint8_t* st = smem + stage * STAGE_BYTES;
        // hi = x rounded to the nearest tf32 (add half an ulp to the bit pattern, clear the 13 low mantissa bits),
        // lo = x - hi (exact in fp32; the tensor core keeps its leading 11 bits).  Pure integer/FADD work -- the
        // conversion pipe (cvt.rna.tf32) is an order of magnitude slower at 16K elements per k-block.
        const uint32_t st_u = smem_u32(st);
#pragma unroll 8
        for (int i = 0; i < 2 * TILE_BYTES / 16 / 128; ++i) {
          const int idx = wtid + i * 128;                         // float4 index over [A tile | B tile]
          const uint32_t q = st_u + (idx < TILE_BYTES / 16 ? 0 : TILE_BYTES) + idx * 16;
          const uint4 v = lds128(q);
          uint4 hi, lo;
          hi.x = (v.x + 0x1000u) & 0xFFFFE000u; hi.y = (v.y + 0x1000u) & 0xFFFFE000u;
          hi.z = (v.z + 0x1000u) & 0xFFFFE000u; hi.w = (v.w + 0x1000u) & 0xFFFFE000u;
          sts128(q, hi);
          if (p.passes == 3) {
            lo.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(hi.x));
            lo.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(hi.y));
            lo.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(hi.z));
            lo.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(hi.w));
            sts128(q + TILE_BYTES, lo);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_split[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (nkb == 0) continue;
      // ---- epilogue for this tile: accumulator row -> registers, 32 columns at a time ----
      mbar_wait(tmem_full, tile & 1);
      tc_fence_after();
      if (wtid == 0 && nt == nt0) DBG_MARK(6);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + c * 32, v);
        if (wtid == 0 && nt == nt0) DBG_MARK(9 + 3 * c);
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) continue;                       // warp-uniform
        if (p.mode != 1) {
          // modes 0 / 2 store a [32 rows x 32 cols] chunk: transpose it through padded shared memory so each
          // store instruction covers 32 consecutive columns of one row (coalesced) instead of 32 different rows
          const uint32_t sc = smem_u32(epi_scratch) + (warp - 2) * (32 * 33 * 4);
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
          if (wtid == 0 && nt == nt0) DBG_MARK(10 + 3 * c);
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
            // all mode / flag decisions are hoisted: each variant is a tight LDS -> (op) -> coalesced STG loop
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
          if (wtid == 0 && nt == nt0) DBG_MARK(11 + 3 * c);
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
      if (wtid == 0 && nt == nt0) DBG_MARK(7);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty);
    }
    if (p.mode == 1 && row_ok) {
      *reinterpret_cast<float4*>(p.part + ((int64_t)blockIdx.y * p.M + row) * 4) = make_float4(rm, rs, rt, rav);
      p.part_idx[(int64_t)blockIdx.y * p.M + row] = rai;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  if (tid == 0) DBG_MARK(8);
}

// ---- host ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map over a row-major [outer, inner] matrix with row stride ld (elements)
static int make_map(CUtensorMap* m, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer,
                    CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = encode_fn();
  DVAE_REQUIRE(fn != nullptr, "tc_gemm: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DVAE_REQUIRE(r == CUDA_SUCCESS, "tc_gemm: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", (int)r,
               (long long)inner, (long long)outer, (long long)ld);
  return DVAE_OK;
}

bool tc_linear_supported(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int K) {
  return ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0 && lda % 4 == 0 && ldb % 4 == 0 &&
         M >= 1 && N >= 1 && K >= 1;
}

static int make_operand_maps(CUtensorMap* ma, CUtensorMap* mb, const float* A, int64_t lda, int trans_a, const float* B,
                             int64_t ldb, int trans_b, int M, int N, int K) {
  int rc;
  // K-major operand: [rows, K] row-major -> inner = K, box 32 x 128, SWIZZLE_128B.
  // MN-major operand: stored [K, rows] -> inner = rows, boxes of 32 x 32, 32-byte-atom swizzle (tf32 requirement).
  if (!trans_a) rc = make_map(ma, A, K, M, lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  else rc = make_map(ma, A, M, K, lda, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  if (!trans_b) rc = make_map(mb, B, K, N, ldb, BK, BN, CU_TENSOR_MAP_SWIZZLE_128B);
  else rc = make_map(mb, B, N, K, ldb, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  return rc;
}

static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, dim3 grid, cudaStream_t st) {
  static bool ready = false;
  if (!ready) {
    DVAE_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    ready = true;
  }
  tc_gemm_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(ma, mb, p);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

__global__ void tc_scale_rows_kernel(float* C, int64_t ldc, int M, int N, float beta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  float* c = C + (i / N) * ldc + (i % N);
  *c = beta == 0.f ? 0.f : *c * beta;
}

int tc_linear_impl(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C, int64_t ldc,
                   int M, int N, int K, const float* bias, const float* bias2, float beta, int act, int passes,
                   cudaStream_t st) {
  DVAE_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "dvae_tc_linear: bad argument");
  DVAE_REQUIRE(passes == 1 || passes == 3, "dvae_tc_linear: passes must be 1 or 3");
  DVAE_REQUIRE(tc_linear_supported(A, lda, B, ldb, M, N, K), "dvae_tc_linear: operands must be 16-byte aligned with ld %% 4 == 0");
  CUtensorMap ma, mb;
  int rc = make_operand_maps(&ma, &mb, A, lda, trans_a, B, ldb, trans_b, M, N, K);
  if (rc) return rc;
  Params p = {};
  p.M = M; p.N = N; p.K = K; p.a_mn = trans_a ? 1 : 0; p.b_mn = trans_b ? 1 : 0; p.passes = passes; p.tiles_per_cta = 1;
  p.C = C; p.ldc = ldc; p.bias = bias; p.bias2 = bias2; p.beta = beta; p.act = act; p.mode = 0;
  if (const char* e = getenv("DVAE_TC_DBG")) p.dbg = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
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
      tc_scale_rows_kernel<<<ceil_div((int64_t)M * N, 256), 256, 0, st>>>(C, ldc, M, N, beta);
      DVAE_LAUNCH_CHECK();
    }
  }
  return launch(ma, mb, p, dim3(ceil_div(M, BM), ceil_div(N, BN), splits), st);
}

// vocab-CE forward: per-(row, vocabulary-split) partials (max, sumexp, target logit, argmax) for rows of h [N,H]
int tc_ce_partials(const float* h, int64_t ldh, int N, int B, int H, int V, const float* w, const float* bias,
                   const int64_t* targets, int64_t tgt_stride_b, const int64_t* lengths, int tiles_per_split, int nsplit,
                   float* part, int* part_idx, const uint64_t* gumbel_seed, uint32_t gumbel_salt, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc = make_operand_maps(&ma, &mb, h, ldh, 0, w, H, 0, N, V, H);
  if (rc) return rc;
  Params p = {};
  p.M = N; p.N = V; p.K = H; p.passes = 3; p.tiles_per_cta = tiles_per_split; p.bias = bias; p.mode = 1;
  p.kb_per_split = ceil_div(H, BK);
  p.targets = targets; p.tgt_stride_b = tgt_stride_b; p.lengths = lengths; p.B = B; p.part = part; p.part_idx = part_idx;
  p.gumbel_seed = gumbel_seed; p.gumbel_salt = gumbel_salt;
  return launch(ma, mb, p, dim3(ceil_div(N, BM), nsplit, 1), st);
}

// softmax-gradient chunk: P[n][v - v0] = (softmax(h W^T + b)[n][v] - [v == target_n]) * mask_n * scale / B, v in [v0, v0+vc)
int tc_softmax_grad(const float* h, int64_t ldh, int N, int B, int H, int v0, int vc, const float* w, const float* bias,
                    const int64_t* targets, int64_t tgt_stride_b, const int64_t* lengths, const float* lse,
                    const float* grad_scale, float* P, int64_t ldp, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc = make_operand_maps(&ma, &mb, h, ldh, 0, w + (int64_t)v0 * H, H, 0, N, vc, H);
  if (rc) return rc;
  Params p = {};
  p.M = N; p.N = vc; p.K = H; p.passes = 3; p.tiles_per_cta = 1; p.bias = bias + v0; p.mode = 2;
  p.kb_per_split = ceil_div(H, BK);
  p.targets = targets; p.tgt_stride_b = tgt_stride_b; p.lengths = lengths; p.B = B; p.lse = lse; p.grad_scale = grad_scale;
  p.v0 = v0; p.C = P; p.ldc = ldp;
  return launch(ma, mb, p, dim3(ceil_div(N, BM), ceil_div(vc, BN), 1), st);
}

}  // namespace tc
}  // namespace dvae

extern "C" int dvae_tc_linear(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C,
                              int64_t ldc, int M, int N, int K, const float* bias, const float* bias2, float beta, int act,
                              int passes, void* stream) {
  return dvae::tc::tc_linear_impl(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, bias2, beta, act, passes,
                                  (cudaStream_t)stream);
}
