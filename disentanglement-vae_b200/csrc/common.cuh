// Shared device/host helpers for the dvae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dvae_b200.h"

namespace dvae {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DVAE_REQUIRE(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::dvae::set_error(__VA_ARGS__);      \
      return DVAE_EINVAL;                  \
    }                                      \
  } while (0)

#define DVAE_CUDA(call)                                                                  \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      ::dvae::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                        cudaGetErrorString(e__));                                        \
      return DVAE_ECUDA;                                                                 \
    }                                                                                    \
  } while (0)

#define DVAE_LAUNCH_CHECK()                                                              \
  do {                                                                                   \
    ::dvae::count_launch();                                                              \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      ::dvae::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,         \
                        cudaGetErrorString(e__));                                        \
      return DVAE_ECUDA;                                                                 \
    }                                                                                    \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Fork / join of independent kernels onto library-owned side streams (also valid inside CUDA graph capture, where it
// becomes parallel branches of the graph).  The small backward GEMMs of one layer are independent of each other and
// each far too small to fill 148 SMs; run back to back they pay their fixed latencies serially.
//   Fork f(st);  launch on st / f.side(0..2);  f.join();     DVAE_FORK=0 keeps everything on the caller's stream.
class Fork {
 public:
  explicit Fork(cudaStream_t main);
  cudaStream_t side(int i);      // i in [0, 3); returns the main stream when forking is disabled
  int join();                    // main waits for every side stream that was used; returns a DVAE_* code
  // Like join(), unless the caller enabled deferred joins (dvae_defer_joins): then the side streams keep running past
  // this call -- e.g. a layer's weight-gradient GEMMs overlap the next layer's recurrence -- until
  // dvae_join_side_streams().  Only for work whose inputs / outputs nothing on the main stream touches before that.
  int join_or_defer();
  // Ordering point for ONE piece of side work that the main stream needs soon while the rest of the fork stays detached:
  // mark(i) right after enqueueing it on side(i), wait_mark() where the main stream consumes it.
  int mark(int i);
  int wait_mark();
  int chain(int from, int to);   // side(to) waits for everything enqueued on side(from) so far
 private:
  cudaStream_t main_;
  bool used_[3] = {false, false, false};
  bool marked_ = false;
  bool ok_;
};

// true between dvae_defer_joins(1) and dvae_join_side_streams(): side work of this call will overlap the caller's next calls
bool defer_joins_enabled();

// DVAE_GEMM_IMPL=simt forces the fp32 SIMT kernels, =tf32 the 3xTF32 tensor-core kernel (A/B tests of the fp16-split path)
bool force_simt_gemm();

// What the caller knows about the dynamic range of a GEMM's operands.  The fp16-split tensor-core kernel multiplies an
// operand by a power of two before splitting it into fp16 planes: either a host constant (`*_scale`) or
// 2^(13 - floor(log2 amax)) from the bit pattern of a device-side max |x| (`*_amax_bits`).  An operand flagged `wide`
// (a gradient of unknown magnitude) without an amax keeps the 3xTF32 kernel, whose operands have fp32 range.
struct GemmHints {
  const uint32_t* a_amax_bits = nullptr;
  const uint32_t* b_amax_bits = nullptr;
  int a_amax_n = 1, b_amax_n = 1;      // number of partial maxima behind each pointer (the kernels take their maximum)
  float a_scale = 1.f, b_scale = 1.f;
  bool a_wide = false, b_wide = false;
  // B operand available as pre-split tile-blocked fp16 planes (tc_planes.cuh, dual-accumulator convention, scale 1): first
  // 128-row block / first k-block of this GEMM inside the planes, and the planes' k-blocks per row block
  const void* b_planes = nullptr;
  int b_tile0 = 0, b_kb0 = 0, b_kbtot = 0;
  int max_ctas = 0;                    // > 0: background GEMM (runs beside a latency-critical kernel): keep the grid at or below
                                       // this many CTAs -- a tensor-core GEMM CTA takes a whole SM (227 KB of shared memory), so a
                                       // machine-filling side GEMM keeps the main stream's next kernel from being scheduled at all
  bool atomic_out = false;             // accumulate into C with atomics even without split-K (C zeroed by the caller): lets two
                                       // GEMMs that add into the same output run concurrently
  bool c_zeroed = false;               // with beta == 0: C already holds zeros (a split-K GEMM then skips its memset node)
  int concurrency = 1;                 // GEMMs of this size the caller runs at the same time (sizes the CTA count to share the SMs)
};
int linear_impl_ex(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C,
                   int64_t ldc, int M, int N, int K, const float* bias, const float* bias2, float beta, int act,
                   const GemmHints& hints, cudaStream_t st);

namespace tc {
bool tc_linear_supported(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int K);
int tc_linear_impl(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C, int64_t ldc,
                   int M, int N, int K, const float* bias, const float* bias2, float beta, int act, int passes,
                   cudaStream_t st);
int tc_ce_partials(const float* h, int64_t ldh, int N, int B, int H, int V, const float* w, const float* bias,
                   const int64_t* targets, int64_t tgt_stride_b, const int64_t* lengths, int tiles_per_split, int nsplit,
                   float* part, int* part_idx, const uint64_t* gumbel_seed, uint32_t gumbel_salt, cudaStream_t st);
int tc_softmax_grad(const float* h, int64_t ldh, int N, int B, int H, int v0, int vc, const float* w, const float* bias,
                    const int64_t* targets, int64_t tgt_stride_b, const int64_t* lengths, const float* lse,
                    const float* grad_scale, float* P, int64_t ldp, cudaStream_t st);
}  // namespace tc

// ---- Philox4x32-10 (counter-based; same mask in forward and backward without storing it) ------
struct Philox {
  static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __host__ __device__ static inline void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(kM0, c[0]), hi1 = __umulhi(kM1, c[2]);
#else
    uint32_t hi0 = (uint32_t)(((uint64_t)kM0 * c[0]) >> 32), hi1 = (uint32_t)(((uint64_t)kM1 * c[2]) >> 32);
#endif
    uint32_t lo0 = kM0 * c[0], lo1 = kM1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  // 4 x uint32 for (seed, salt, 64-bit counter)
  __host__ __device__ static inline void gen(uint64_t seed, uint32_t salt, uint64_t ctr, uint32_t (&out)[4]) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), salt, 0x5eedu};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += kW0;
      k1 += kW1;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

// Dropout keep-scale for element `idx` (idx = row * width + col): 0 or 1/(1-p).
// One Philox call serves 4 consecutive elements (idx / 4 is the counter, idx % 4 the lane).
__device__ __forceinline__ void dropout_scale4(uint64_t seed, uint32_t salt, uint64_t idx4, float p,
                                               float inv_keep, float (&s)[4]) {
  uint32_t r[4];
  Philox::gen(seed, salt, idx4, r);
  // keep iff u >= p with u = r * 2^-32 in [0,1)
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i] = ((float)r[i] * 2.3283064365386963e-10f >= p) ? inv_keep : 0.f;
}

// Gumbel(0,1) noise for (row, col..col+3): one Philox call per 4 consecutive columns (col % 4 == 0).
// u is a 24-bit uniform in (0,1) so the host can reproduce it exactly.  argmax_v(logit_v + g_v) ~ softmax(logits).
__device__ __forceinline__ void gumbel4(uint64_t seed, uint32_t salt, int row, int col, int ncol4, float (&g)[4]) {
  uint32_t r[4];
  Philox::gen(seed, salt, (uint64_t)row * ncol4 + (col >> 2), r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float u = ((float)(r[i] >> 8) + 0.5f) * (1.f / 16777216.f);
    g[i] = -logf(-logf(u));
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

}  // namespace dvae
