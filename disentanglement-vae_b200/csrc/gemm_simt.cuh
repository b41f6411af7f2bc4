// fp32 SIMT tile GEMM main loop, shared by dvae_linear and the vocabulary-CE kernels.
//
// acc[TM][TN] = sum_k A(m,k) * B(n,k) for one BM x BN tile; operands may be K-contiguous
// (row-major [rows,K]) or row-contiguous (stored [K,rows]).  Global -> register -> shared
// double buffering; shared tiles are K-major ([BK][rows+4]) so the inner product reads float4s.
// Exact fp32 FFMA accumulation: this is the parity path (fp32 loss within 1e-5, argmax identical).
#pragma once
#include "common.cuh"

namespace dvae {

template <int BM_, int BN_, int BK_, int TM_, int TN_, bool A_KC, bool B_KC>
struct GemmTile {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, TM = TM_, TN = TN_;
  static constexpr int TX = BN / TN, TY = BM / TM, NT = TX * TY;
  static constexpr int PAD = 4;
  static constexpr int LDA_S = BM + PAD, LDB_S = BN + PAD;
  static constexpr int STAGE_FLOATS = BK * (LDA_S + LDB_S);
  static constexpr int SMEM_FLOATS = 2 * STAGE_FLOATS;
  static constexpr int A_VECS = BM * BK / 4 / NT, B_VECS = BN * BK / 4 / NT;
  static_assert(BM * BK / 4 % NT == 0 && BN * BK / 4 % NT == 0, "tile/threads mismatch");
  static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile must be a multiple of 4");

  // row / column owned by this thread for micro-tile index i (interleaved halves: conflict-free LDS.128)
  __device__ static __forceinline__ int row_of(int ty, int i) { return (i / 4) * (BM / (TM / 4)) + ty * 4 + (i % 4); }
  __device__ static __forceinline__ int col_of(int tx, int j) { return (j / 4) * (BN / (TN / 4)) + tx * 4 + (j % 4); }

  template <int BR, int VECS, bool KC>
  __device__ static __forceinline__ void gload(const float* __restrict__ p, int64_t ld, bool vec_ok, int row0,
                                               int nrows, int k0, int K, int tid, float4 (&st)[VECS]) {
#pragma unroll
    for (int i = 0; i < VECS; ++i) {
      int v = i * NT + tid;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (KC) {
        int r = v / (BK / 4), kq = v % (BK / 4);
        int gr = row0 + r, gk = k0 + kq * 4;
        if (gr < nrows) {
          const float* q = p + (int64_t)gr * ld + gk;
          if (vec_ok && gk + 3 < K) {
            val = *reinterpret_cast<const float4*>(q);
          } else {
            if (gk + 0 < K) val.x = q[0];
            if (gk + 1 < K) val.y = q[1];
            if (gk + 2 < K) val.z = q[2];
            if (gk + 3 < K) val.w = q[3];
          }
        }
      } else {
        int k = v / (BR / 4), rq = v % (BR / 4);
        int gk = k0 + k, gr = row0 + rq * 4;
        if (gk < K) {
          const float* q = p + (int64_t)gk * ld + gr;
          if (vec_ok && gr + 3 < nrows) {
            val = *reinterpret_cast<const float4*>(q);
          } else {
            if (gr + 0 < nrows) val.x = q[0];
            if (gr + 1 < nrows) val.y = q[1];
            if (gr + 2 < nrows) val.z = q[2];
            if (gr + 3 < nrows) val.w = q[3];
          }
        }
      }
      st[i] = val;
    }
  }

  template <int BR, int LDS_, int VECS, bool KC>
  __device__ static __forceinline__ void sstore(float* __restrict__ s, int tid, const float4 (&st)[VECS]) {
#pragma unroll
    for (int i = 0; i < VECS; ++i) {
      int v = i * NT + tid;
      if (KC) {
        int r = v / (BK / 4), kq = v % (BK / 4);
        s[(kq * 4 + 0) * LDS_ + r] = st[i].x;
        s[(kq * 4 + 1) * LDS_ + r] = st[i].y;
        s[(kq * 4 + 2) * LDS_ + r] = st[i].z;
        s[(kq * 4 + 3) * LDS_ + r] = st[i].w;
      } else {
        int k = v / (BR / 4), rq = v % (BR / 4);
        *reinterpret_cast<float4*>(&s[k * LDS_ + rq * 4]) = st[i];
      }
    }
  }

  // Accumulates into acc over k in [k_begin, k_end).  smem: SMEM_FLOATS floats, 16-byte aligned.
  __device__ static __forceinline__ void run(const float* __restrict__ A, int64_t lda, int m0, int M,
                                             const float* __restrict__ B, int64_t ldb, int n0, int N,
                                             int k_begin, int k_end, int K, float* smem,
                                             float (&acc)[TM][TN]) {
    const int tid = threadIdx.x, ty = tid / TX, tx = tid % TX;
    const bool a_vec = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 4 == 0);
    const bool b_vec = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (ldb % 4 == 0);
    float4 sta[A_VECS], stb[B_VECS];
    float* As = smem;
    float* Bs = smem + BK * LDA_S;
    const int nk = (k_end - k_begin + BK - 1) / BK;
    if (nk <= 0) return;
    gload<BM, A_VECS, A_KC>(A, lda, a_vec, m0, M, k_begin, min(K, k_end), tid, sta);
    gload<BN, B_VECS, B_KC>(B, ldb, b_vec, n0, N, k_begin, min(K, k_end), tid, stb);
    sstore<BM, LDA_S, A_VECS, A_KC>(As, tid, sta);
    sstore<BN, LDB_S, B_VECS, B_KC>(Bs, tid, stb);
    __syncthreads();
    for (int it = 0; it < nk; ++it) {
      float* Ac = smem + (it & 1) * STAGE_FLOATS;
      float* Bc = Ac + BK * LDA_S;
      if (it + 1 < nk) {
        int k0 = k_begin + (it + 1) * BK;
        gload<BM, A_VECS, A_KC>(A, lda, a_vec, m0, M, k0, min(K, k_end), tid, sta);
        gload<BN, B_VECS, B_KC>(B, ldb, b_vec, n0, N, k0, min(K, k_end), tid, stb);
      }
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM / 4; ++i) {
          float4 t = *reinterpret_cast<const float4*>(&Ac[k * LDA_S + i * (BM / (TM / 4)) + ty * 4]);
          a[i * 4 + 0] = t.x; a[i * 4 + 1] = t.y; a[i * 4 + 2] = t.z; a[i * 4 + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < TN / 4; ++j) {
          float4 t = *reinterpret_cast<const float4*>(&Bc[k * LDB_S + j * (BN / (TN / 4)) + tx * 4]);
          b[j * 4 + 0] = t.x; b[j * 4 + 1] = t.y; b[j * 4 + 2] = t.z; b[j * 4 + 3] = t.w;
        }
        // packed fp32 FMA: (acc[i][j], acc[i][j+1]) += (a[i], a[i]) * (b[j], b[j+1]).  Two IEEE fmas per
        // FFMA2, bit-identical to the scalar form; scalar FFMA issues at half rate on sm_100.
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float2 aa = make_float2(a[i], a[i]);
#pragma unroll
          for (int j = 0; j < TN; j += 2) {
            float2 c = make_float2(acc[i][j], acc[i][j + 1]);
            c = __ffma2_rn(aa, make_float2(b[j], b[j + 1]), c);
            acc[i][j] = c.x; acc[i][j + 1] = c.y;
          }
        }
      }
      if (it + 1 < nk) {
        float* An = smem + ((it + 1) & 1) * STAGE_FLOATS;
        sstore<BM, LDA_S, A_VECS, A_KC>(An, tid, sta);
        sstore<BN, LDB_S, B_VECS, B_KC>(An + BK * LDA_S, tid, stb);
      }
      __syncthreads();
    }
  }
};

}  // namespace dvae
