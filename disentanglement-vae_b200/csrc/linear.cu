// dvae_linear / dvae_colsum: dense fp32 layers on the path (see include/dvae_b200.h).
#include "gemm_simt.cuh"
#include "tc_gemm16.cuh"

namespace dvae {

template <class G>
__global__ void __launch_bounds__(G::NT) linear_kernel(const float* __restrict__ A, int64_t lda,
                                                        const float* __restrict__ B, int64_t ldb,
                                                        float* __restrict__ C, int64_t ldc, int M, int N, int K,
                                                        const float* __restrict__ bias,
                                                        const float* __restrict__ bias2, float beta, int act,
                                                        int k_per_split) {
  __shared__ __align__(16) float smem[G::SMEM_FLOATS];
  const int m0 = blockIdx.y * G::BM, n0 = blockIdx.x * G::BN;
  const int kb = blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
  float acc[G::TM][G::TN];
#pragma unroll
  for (int i = 0; i < G::TM; ++i)
#pragma unroll
    for (int j = 0; j < G::TN; ++j) acc[i][j] = 0.f;
  G::run(A, lda, m0, M, B, ldb, n0, N, kb, ke, K, smem, acc);
  const int ty = threadIdx.x / G::TX, tx = threadIdx.x % G::TX;
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < G::TM; ++i) {
    int m = m0 + G::row_of(ty, i);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < G::TN; ++j) {
      int n = n0 + G::col_of(tx, j);
      if (n >= N) continue;
      float v = acc[i][j];
      float* c = C + (int64_t)m * ldc + n;
      if (split) {
        // C was pre-scaled by beta on the host side (memset for beta == 0); bias once, by split 0
        if (blockIdx.z == 0) {
          if (bias) v += bias[n];
          if (bias2) v += bias2[n];
        }
        atomicAdd(c, v);
      } else {
        if (bias) v += bias[n];
        if (bias2) v += bias2[n];
        if (act == 1) v = tanhf(v);
        if (beta != 0.f) v += beta * (*c);
        *c = v;
      }
    }
  }
}

using GBig_NT = GemmTile<128, 128, 16, 8, 8, true, true>;
using GBig_NN = GemmTile<128, 128, 16, 8, 8, true, false>;
using GBig_TN = GemmTile<128, 128, 16, 8, 8, false, false>;
using GBig_TT = GemmTile<128, 128, 16, 8, 8, false, true>;
using GSm_NT = GemmTile<64, 64, 16, 4, 4, true, true>;
using GSm_NN = GemmTile<64, 64, 16, 4, 4, true, false>;
using GSm_TN = GemmTile<64, 64, 16, 4, 4, false, false>;
using GSm_TT = GemmTile<64, 64, 16, 4, 4, false, true>;

template <class G>
static int launch_linear(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M,
                         int N, int K, const float* bias, const float* bias2, float beta, int act, int splits,
                         cudaStream_t st) {
  int k_per_split = ceil_div(ceil_div(K, splits), G::BK) * G::BK;
  splits = ceil_div(K, k_per_split);
  dim3 grid(ceil_div(N, G::BN), ceil_div(M, G::BM), splits);
  linear_kernel<G><<<grid, G::NT, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, bias2, beta, act, k_per_split);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

__global__ void scale_rows_kernel(float* C, int64_t ldc, int M, int N, float beta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  float* c = C + (i / N) * ldc + (i % N);
  *c = beta == 0.f ? 0.f : *c * beta;
}

int linear_impl(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C,
                int64_t ldc, int M, int N, int K, const float* bias, const float* bias2, float beta, int act,
                cudaStream_t st) {
  return linear_impl_ex(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, bias2, beta, act, GemmHints(), st);
}

int linear_impl_ex(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C,
                   int64_t ldc, int M, int N, int K, const float* bias, const float* bias2, float beta, int act,
                   const GemmHints& hints, cudaStream_t st) {
  DVAE_REQUIRE(A && B && C, "dvae_linear: null pointer");
  DVAE_REQUIRE(M > 0 && N > 0 && K > 0, "dvae_linear: non-positive size M=%d N=%d K=%d", M, N, K);
  DVAE_REQUIRE(act == 0 || act == 1, "dvae_linear: unknown activation %d", act);
  // Dense-contraction shapes go to the tensor cores (TMA + tcgen05); small (fixed costs of the tensor-core kernels:
  // ~10 us) or unaligned problems stay on the fp32 SIMT kernels below.
  static const double tc_min_mnk = [] {
    const char* e = getenv("DVAE_TC_MIN_LOG2_MNK");      // A/B knob: smallest M*N*K (log2) that goes to the tensor-core kernels
    return (double)(1ll << (e ? atoi(e) : 24));      // measured on the cfg-2 step: 23: 1.455, 24: 1.438, 25: 1.449, 26: 1.453 ms
  }();
  if (!force_simt_gemm() && M >= 64 && N >= 64 && K >= 32 && (double)M * N * K >= tc_min_mnk) {
    // fp16-split kernel unless an operand's dynamic range is unknown (then 3xTF32, whose operands have fp32 range)
    const bool unscaled_wide = (hints.a_wide && !hints.a_amax_bits) || (hints.b_wide && !hints.b_amax_bits);
    if (!unscaled_wide && tc16::supported(A, lda, trans_a, B, ldb, trans_b, M, N, K))
      return tc16::linear(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, bias2, beta, act, hints, st);
    if (tc::tc_linear_supported(A, lda, B, ldb, M, N, K))
      return tc::tc_linear_impl(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, bias2, beta, act, 3, st);
  }
  // Tile choice: 128x128 when that already fills the 148 SMs, else 64x64; split-K (atomic
  // accumulation) when even the small tiles leave most SMs idle and K is deep.
  const int kSMs = 148;
  int64_t big_tiles = (int64_t)ceil_div(M, 128) * ceil_div(N, 128);
  int64_t small_tiles = (int64_t)ceil_div(M, 64) * ceil_div(N, 64);
  bool big = big_tiles >= kSMs;
  int splits = 1;
  if (!big && small_tiles < kSMs / 2 && K >= 512 && act == 0) {
    { int64_t want = (kSMs + small_tiles - 1) / small_tiles, cap = K / 128; splits = (int)(want < cap ? want : cap); }
    if (splits < 1) splits = 1;
  }
  if (splits > 1) {
    // pre-scale C by beta so the splits can accumulate atomically
    if (beta != 1.f) {
      int64_t n = (int64_t)M * N;
      scale_rows_kernel<<<ceil_div(n, 256), 256, 0, st>>>(C, ldc, M, N, beta);
      DVAE_LAUNCH_CHECK();
    }
  }
#define DVAE_DISPATCH(GB, GS)                                                                                    \
  return big ? launch_linear<GB>(A, lda, B, ldb, C, ldc, M, N, K, bias, bias2, beta, act, 1, st)                 \
             : launch_linear<GS>(A, lda, B, ldb, C, ldc, M, N, K, bias, bias2, beta, act, splits, st)
  if (!trans_a && !trans_b) { DVAE_DISPATCH(GBig_NT, GSm_NT); }
  if (!trans_a && trans_b) { DVAE_DISPATCH(GBig_NN, GSm_NN); }
  if (trans_a && trans_b) { DVAE_DISPATCH(GBig_TN, GSm_TN); }
  DVAE_DISPATCH(GBig_TT, GSm_TT);
#undef DVAE_DISPATCH
}

// out[n] (+)= sum_m X[m][n]: blocks of 32 columns x a slice of the rows (grid.y), 32x8 threads, coalesced row
// reads; row slices combine with one atomicAdd per (column, slice) into the pre-scaled output.
__global__ void colsum_kernel(const float* __restrict__ X, int64_t ldx, int M, int N, float* __restrict__ out,
                              int rows_per_block) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f;
  if (n < N) {
    int m = m0 + threadIdx.y;
    for (; m + 8 < m1; m += 16) { s0 += X[(int64_t)m * ldx + n]; s1 += X[(int64_t)(m + 8) * ldx + n]; }
    if (m < m1) s0 += X[(int64_t)m * ldx + n];
  }
  red[threadIdx.y][threadIdx.x] = s0 + s1;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + n, t);
  }
}

__global__ void scale_vec_kernel(float* out, int N, float beta) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) out[i] *= beta;
}

int colsum_impl(const float* X, int64_t ldx, int M, int N, float* out, float beta, cudaStream_t st) {
  DVAE_REQUIRE(X && out && M > 0 && N > 0, "dvae_colsum: bad argument");
  if (beta == 0.f) {
    DVAE_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, st));
  } else if (beta != 1.f) {
    scale_vec_kernel<<<ceil_div(N, 256), 256, 0, st>>>(out, N, beta);
    DVAE_LAUNCH_CHECK();
  }
  const int col_blocks = ceil_div(N, 32);
  int row_blocks = ceil_div(2 * 148, col_blocks);                 // aim for ~2 CTAs per SM
  if (row_blocks > ceil_div(M, 64)) row_blocks = ceil_div(M, 64);
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = ceil_div(M, row_blocks);
  colsum_kernel<<<dim3(col_blocks, ceil_div(M, rows_per_block)), dim3(32, 8), 0, st>>>(X, ldx, M, N, out, rows_per_block);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

}  // namespace dvae

extern "C" int dvae_linear(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b,
                           float* C, int64_t ldc, int M, int N, int K, const float* bias, const float* bias2,
                           float beta, int act, void* stream) {
  dvae::GemmHints hints;           // a public entry point knows nothing about its operands' dynamic range
  hints.a_wide = hints.b_wide = true;
  return dvae::linear_impl_ex(A, lda, trans_a, B, ldb, trans_b, C, ldc, M, N, K, bias, bias2, beta, act, hints,
                           (cudaStream_t)stream);
}

extern "C" int dvae_colsum(const float* X, int64_t ldx, int M, int N, float* out, float beta, void* stream) {
  return dvae::colsum_impl(X, ldx, M, N, out, beta, (cudaStream_t)stream);
}
