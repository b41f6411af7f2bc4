// Probe: cycles per tcgen05.mma.kind::f16 (M=128, K=16) at small N, as a function of
//   - operand source for A: shared memory (SS) vs tensor memory (TS)
//   - number of independent accumulators the issue loop round-robins over (dependent-accumulate latency)
//   - N (16, 32, 64)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I disentanglement-vae_b200/csrc \
//        profiles/probes/umma_small_n.cu -o profiles/probes/bin/umma_small_n
#include <cstdio>
#include <cuda_fp16.h>
#include "tc_gemm.cuh"
namespace dvae { void set_error(const char*, ...) {} void count_launch(int) {} }
using namespace dvae::tc;

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

// issue-style probe: 48 MMAs (16 k-steps x 3 products, as the LSTM step does), fully unrolled, N = 16
template <int STYLE>
__device__ __forceinline__ void issue48(uint32_t tmem, uint32_t sb, uint64_t* bar) {
  constexpr uint32_t idesc = idesc_f16(128, 16);
  const uint64_t ahi = make_smem_desc(sb, 128, 4096, 0), alo = make_smem_desc(sb + 65536, 128, 4096, 0);
  const uint64_t bhi = make_smem_desc(sb + 131072, 256, 128, 0), blo = make_smem_desc(sb + 131072 + 8192, 256, 128, 0);
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) {
    const uint64_t da = (uint64_t)(ks * 16), db = (uint64_t)(ks * 32);
    mma_ss(tmem, ahi + da, bhi + db, idesc, ks > 0);
    mma_ss(tmem + 16, ahi + da, blo + db, idesc, ks > 0);
    mma_ss(tmem + 16, alo + da, bhi + db, idesc, 1);
  }
  tc_commit(bar);
}

// layout probe: 48 MMAs M=128 N=128 K=16 (8 k-steps x 3 products x 2), no-swizzle K-major operands, two core-matrix placements
template <int LBO, int SBO, int KSTEP>
__device__ __forceinline__ void issue_n128(uint32_t tmem, uint32_t sb, uint64_t* bar) {
  constexpr uint32_t idesc = idesc_f16(128, 128);
  const uint64_t ahi = make_smem_desc(sb, LBO, SBO, 0), alo = make_smem_desc(sb + 16384, LBO, SBO, 0);
  const uint64_t bhi = make_smem_desc(sb + 32768, LBO, SBO, 0), blo = make_smem_desc(sb + 49152, LBO, SBO, 0);
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) {
    const uint64_t d = (uint64_t)(((ks & 1) * KSTEP) >> 4);     // the two k-steps of one 32-k block, re-read (timing only)
    mma_ss(tmem, ahi + d, bhi + d, idesc, ks > 0);
    mma_ss(tmem + 128, ahi + d, blo + d, idesc, ks > 0);
    mma_ss(tmem + 128, alo + d, bhi + d, idesc, 1);
  }
  tc_commit(bar);
}

// commit-frequency probe: the same 48 MMAs (M=128 N=128), with a tcgen05.commit after every `per` MMAs
template <int PER>
__device__ __forceinline__ void issue_commit_every(uint32_t tmem, uint32_t sb, uint64_t* bar, uint64_t* dummy) {
  constexpr uint32_t idesc = idesc_f16(128, 128);
  const uint64_t ahi = make_smem_desc(sb, 128, 512, 0), alo = make_smem_desc(sb + 16384, 128, 512, 0);
  const uint64_t bhi = make_smem_desc(sb + 32768, 128, 512, 0), blo = make_smem_desc(sb + 49152, 128, 512, 0);
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) {
    const uint64_t d = (uint64_t)(((ks & 1) * 256) >> 4);
    mma_ss(tmem, ahi + d, bhi + d, idesc, ks > 0);
    mma_ss(tmem + 128, ahi + d, blo + d, idesc, ks > 0);
    mma_ss(tmem + 128, alo + d, bhi + d, idesc, 1);
    if (PER > 0 && ((ks + 1) * 3) % PER == 0 && ks != 15) tc_commit(dummy);
  }
  tc_commit(bar);
}

__global__ void __launch_bounds__(128, 1) probe_commit(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, dummy;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&dummy, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, sb = smem_u32(smem);
  int phase = 0;
  for (int rep = 0; rep < 2; ++rep) {
    for (int v = 0; v < 3; ++v) {
      long long t0 = clock64(), t1 = 0;
      if (warp == 0) {
        if (elect_one()) {
          if (v == 0) issue_commit_every<0>(tmem, sb, &bar, &dummy);
          else if (v == 1) issue_commit_every<12>(tmem, sb, &bar, &dummy);
          else issue_commit_every<6>(tmem, sb, &bar, &dummy);
          t1 = clock64();
        }
        __syncwarp();
      }
      mbar_wait(&bar, phase); phase ^= 1;
      long long t2 = clock64();
      if (tid == 0) { out[rep * 6 + 2 * v] = t1 - t0; out[rep * 6 + 2 * v + 1] = t2 - t0; }
      __syncthreads();
    }
  }
  if (tid < 32) tmem_dealloc(tmem, 256);
}

__global__ void __launch_bounds__(128, 1) probe_layout(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, sb = smem_u32(smem);
  int phase = 0;
  for (int rep = 0; rep < 2; ++rep) {
    for (int lay = 0; lay < 3; ++lay) {
      long long t0 = clock64();
      if (warp == 0) {
        if (elect_one()) {
          if (lay == 0) issue_n128<128, 512, 256>(tmem, sb, &bar);          // row-group blocks of 4 chunks (dense, tc_gemm16 pre-split)
          else if (lay == 1) issue_n128<160, 640, 320>(tmem, sb, &bar);     // padded (tc_gemm16 converter path)
          else issue_n128<2048, 128, 4096>(tmem, sb, &bar);                 // k-chunk columns of 128 rows x 16 B
        }
        __syncwarp();
      }
      mbar_wait(&bar, phase); phase ^= 1;
      long long t2 = clock64();
      if (tid == 0) out[rep * 3 + lay] = t2 - t0;
      __syncthreads();
    }
  }
  if (tid < 32) tmem_dealloc(tmem, 256);
}

__global__ void __launch_bounds__(128, 1) probe_style(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, sb = smem_u32(smem);
  int phase = 0;
  for (int rep = 0; rep < 3; ++rep) {
    // style 0: divergent single thread
    long long t0 = clock64(), t1 = 0;
    if (tid == 0) { issue48<0>(tmem, sb, &bar); t1 = clock64(); }
    mbar_wait(&bar, phase); phase ^= 1;
    long long t2 = clock64();
    if (tid == 0) { out[rep * 4 + 0] = t1 - t0; out[rep * 4 + 1] = t2 - t0; }
    __syncthreads();
    // style 1: warp-uniform branch + elect.sync
    t0 = clock64(); t1 = 0;
    if (warp == 0) {
      if (elect_one()) { issue48<1>(tmem, sb, &bar); t1 = clock64(); }
      __syncwarp();
    }
    mbar_wait(&bar, phase); phase ^= 1;
    t2 = clock64();
    if (tid == 0) { out[rep * 4 + 2] = t1 - t0; out[rep * 4 + 3] = t2 - t0; }
    __syncthreads();
  }
  if (tid < 32) tmem_dealloc(tmem, 64);
}

__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, sb = smem_u32(smem);
  if (tid == 0) {
    int phase = 0, o = 0;
    const int Ns[3] = {16, 32, 64};
    for (int ni = 0; ni < 3; ++ni) {
      const int N = Ns[ni];
      const uint32_t idesc = idesc_f16(128, N);
      for (int mode = 0; mode < 2; ++mode) {          // 0 = SS, 1 = TS
        for (int nacc = 1; nacc <= 4; ++nacc) {
          const int R = 96;
          long long t0 = clock64();
          for (int r = 0; r < R; ++r) {
            const int ks = r & 15;
            const uint64_t a = make_smem_desc(sb + ks * 256, 128, 4096, 0);
            const uint64_t b = make_smem_desc(sb + 131072 + ks * 2 * N * 16, N * 16, 128, 0);
            const uint32_t d = tmem + 256 + (r % nacc) * N;
            if (mode == 0) mma_ss(d, a, b, idesc, r >= nacc);
            else mma_ts(d, tmem + ks * 8, b, idesc, r >= nacc);
          }
          tc_commit(&bar);
          long long t1 = clock64();
          mbar_wait(&bar, phase); phase ^= 1;
          long long t2 = clock64();
          out[o++] = t1 - t0; out[o++] = t2 - t0;
        }
      }
    }
  }
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}


// contention probe: does a stream of M=128 N=128 MMAs (SS: 8 KB of SMEM operand reads per 64-cycle MMA) slow down when
// the async proxy writes operand tiles into the same SM's shared memory at the same time (and vice versa)?
//   mode bit0: MMA chain (48 x R MMAs), bit1: bulk-copy stream (C copies of 16 KB into a 4-slot ring), bit2: A from TMEM
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(smem_u32(bar)) : "memory");
}
//   mode bit3: random fp16 operand bits instead of 1.0, bit4: warps 2.. spin on an mbarrier for the whole run
__global__ void __launch_bounds__(832, 1) probe_contend(long long* out, const uint8_t* src, int mode, int R, int C) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, full[4], fin;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;
    if (mode & 8) {                      // two random fp16 in [-2, 2): sign + exponent 0x3c/0x38.. + random mantissa
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x83ff83ffu) | 0x38003800u | ((h >> 3) & 0x04000400u);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&fin, 2); for (int i = 0; i < 4; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, sb = smem_u32(smem);
  const long long t0 = clock64();
  if (warp == 0 && (mode & 1)) {
    if (elect_one()) {
      constexpr uint32_t idesc = idesc_f16(128, 128);
      const uint64_t ahi = make_smem_desc(sb, 128, 512, 0), alo = make_smem_desc(sb + 16384, 128, 512, 0);
      const uint64_t bhi = make_smem_desc(sb + 32768, 128, 512, 0), blo = make_smem_desc(sb + 49152, 128, 512, 0);
      for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint64_t d = (uint64_t)(((ks & 1) * 256) >> 4);
          if (mode & 4) {
            mma_ts(tmem, tmem + 256 + (ks & 1) * 8, bhi + d, idesc, 1);
            mma_ts(tmem + 128, tmem + 256 + (ks & 1) * 8, blo + d, idesc, 1);
            mma_ts(tmem + 128, tmem + 384 + (ks & 1) * 8, bhi + d, idesc, 1);
          } else {
            mma_ss(tmem, ahi + d, bhi + d, idesc, 1);
            mma_ss(tmem + 128, ahi + d, blo + d, idesc, 1);
            mma_ss(tmem + 128, alo + d, bhi + d, idesc, 1);
          }
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (tid == 0) out[blockIdx.x * 2] = clock64() - t0;
  }
  if (warp == 0 && tid == 0) mbar_arrive(&fin);
  if (warp >= 2 && (mode & 16)) mbar_wait(&fin, 0);
  if (warp == 1 && (mode & 2)) {
    if (elect_one()) {
      const uint8_t* g = src + (size_t)blockIdx.x * 16384;
      for (int i = 0; i < C; ++i) {
        const int s = i & 3;
        if (i >= 4) mbar_wait(&full[s], ((i >> 2) - 1) & 1);
        mbar_expect_tx(&full[s], 16384);
        bulk_g2s(sb + 65536 + s * 16384, g + (size_t)(i & 7) * (148 * 16384), 16384, &full[s]);
      }
      for (int i = C; i < C + 4; ++i) mbar_wait(&full[i & 3], ((i >> 2) - 1) & 1);
      out[blockIdx.x * 2 + 1] = clock64() - t0;
    }
    __syncwarp();
  }
  if (warp == 1 && (tid & 31) == 0) mbar_arrive(&fin);
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}


// loop-form probe: the tc_gemm16 main loop without any waiting -- 16 tiles x 8 k-blocks x (2 k-steps x 3 MMAs), operands at
// runtime-computed ring addresses.  V0: elect inside the k-block loop (as the kernel does), V1: one elected thread runs the
// whole loop, V2: V1 + all ring descriptors precomputed before the loop (no descriptor arithmetic between MMAs),
// V3: V1 with the k-block loop unrolled by 4
template <int V>
__global__ void __launch_bounds__(128, 1) probe_loop(long long* out, int ntiles, int nkb, int nstages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 216 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, smem_u = smem_u32(smem);
  constexpr uint32_t idesc = idesc_f16(128, 128);
  constexpr uint32_t PLANE = 8192, A_BYTES = 131072, STAGE = 16384;
  const long long t0 = clock64();
  if (warp == 0) {
    auto body = [&](int tile, int kb, int stage) {
      const uint32_t d1 = tmem + (tile & 1) * 256, d2 = d1 + 128;
      const uint32_t sa = smem_u + kb * 2 * PLANE, sb = smem_u + A_BYTES + stage * STAGE;
      const uint64_t ahi = make_smem_desc(sa, 128, 512, 0), alo = make_smem_desc(sa + PLANE, 128, 512, 0);
      const uint64_t bhi = make_smem_desc(sb, 128, 512, 0), blo = make_smem_desc(sb + PLANE, 128, 512, 0);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint64_t adv = (uint64_t)(k * 256 >> 4);
        const uint32_t accum = (kb | k) ? 1u : 0u;
        mma_ss(d1, ahi + adv, bhi + adv, idesc, accum);
        mma_ss(d2, ahi + adv, blo + adv, idesc, accum);
        mma_ss(d2, alo + adv, bhi + adv, idesc, 1u);
      }
    };
    if (V == 0) {
      int stage = 0;
      for (int tile = 0; tile < ntiles; ++tile)
        for (int kb = 0; kb < nkb; ++kb) {
          if (elect_one()) body(tile, kb, stage);
          __syncwarp();
          if (++stage == nstages) stage = 0;
        }
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
    } else if (elect_one()) {
      if (V == 1) {
        int stage = 0;
        for (int tile = 0; tile < ntiles; ++tile)
          for (int kb = 0; kb < nkb; ++kb) {
            body(tile, kb, stage);
            if (++stage == nstages) stage = 0;
          }
      } else if (V == 3) {
        int stage = 0;
        for (int tile = 0; tile < ntiles; ++tile) {
#pragma unroll 4
          for (int kb = 0; kb < nkb; ++kb) {
            body(tile, kb, stage);
            if (++stage == nstages) stage = 0;
          }
        }
      } else {
        // descriptors for the 8 stationary A k-blocks and up to 5 B stages, kept in registers
        uint64_t da[8][2], db[5][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) { da[i][0] = make_smem_desc(smem_u + i * 2 * PLANE, 128, 512, 0); da[i][1] = make_smem_desc(smem_u + i * 2 * PLANE + PLANE, 128, 512, 0); }
#pragma unroll
        for (int i = 0; i < 5; ++i) { db[i][0] = make_smem_desc(smem_u + A_BYTES + i * STAGE, 128, 512, 0); db[i][1] = make_smem_desc(smem_u + A_BYTES + i * STAGE + PLANE, 128, 512, 0); }
        // 40 k-blocks = lcm(8, 5): fully unrolled so that every descriptor is a compile-time register choice
        for (int rep = 0; rep < ntiles * nkb / 40; ++rep) {
#pragma unroll
          for (int i = 0; i < 40; ++i) {
            const int kb = i & 7, stage = i % 5, tile = i >> 3;
            const uint32_t d1 = tmem + (tile & 1) * 256, d2 = d1 + 128;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const uint64_t adv = (uint64_t)(k * 256 >> 4);
              mma_ss(d1, da[kb][0] + adv, db[stage][0] + adv, idesc, (kb | k) ? 1u : 0u);
              mma_ss(d2, da[kb][0] + adv, db[stage][1] + adv, idesc, (kb | k) ? 1u : 0u);
              mma_ss(d2, da[kb][1] + adv, db[stage][0] + adv, idesc, 1u);
            }
          }
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  if (tid == 0) out[0] = clock64() - t0;
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}


// two-issuer probe: the LSTM step's 48 small MMAs (M=128, N=16, K=16; 16 k-steps x 3 products) issued by ONE thread, or
// split between TWO warps (k-steps 0-7 / 8-15, separate accumulators): is the ~36 cycles per small MMA a per-thread
// issue cost or a limit of the tensor-core front end?
template <int K0, int K1>
__device__ __forceinline__ void issue_ksteps(uint32_t tmem, uint32_t sb, uint64_t* bar) {
  constexpr uint32_t idesc = idesc_f16(128, 16);
  const uint64_t ahi = make_smem_desc(sb, 128, 4096, 0), alo = make_smem_desc(sb + 65536, 128, 4096, 0);
  const uint64_t bhi = make_smem_desc(sb + 131072, 256, 128, 0), blo = make_smem_desc(sb + 131072 + 8192, 256, 128, 0);
#pragma unroll
  for (int ks = K0; ks < K1; ++ks) {
    const uint64_t da = (uint64_t)(ks * 16), db = (uint64_t)(ks * 32);
    mma_ss(tmem, ahi + da, bhi + db, idesc, ks > K0);
    mma_ss(tmem + 16, ahi + da, blo + db, idesc, ks > K0);
    mma_ss(tmem + 16, alo + da, bhi + db, idesc, 1);
  }
  tc_commit(bar);
}
__global__ void __launch_bounds__(128, 1) probe_two_issuers(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 2); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot, sb = smem_u32(smem);
  for (int rep = 0; rep < 3; ++rep) {
    long long t0 = clock64();
    if (warp == 0) {
      if (elect_one()) issue_ksteps<0, 16>(tmem, sb, &bar);
      __syncwarp();
    }
    mbar_wait(&bar, rep & 1);
    long long t1 = clock64();
    __syncthreads();
    long long t2 = clock64();
    if (warp == 0) {
      if (elect_one()) issue_ksteps<0, 8>(tmem, sb, &bar2);
      __syncwarp();
    } else if (warp == 1) {
      if (elect_one()) issue_ksteps<8, 16>(tmem + 32, sb, &bar2);
      __syncwarp();
    }
    mbar_wait(&bar2, rep & 1);
    long long t3 = clock64();
    if (tid == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t3 - t2; }
    __syncthreads();
  }
  if (tid < 32) tmem_dealloc(tmem, 128);
}

int main() {
  long long* d; cudaMalloc(&d, 1024);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe_style, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int it = 0; it < 2; ++it) probe<<<1, 128, 200 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  long long h[128]; cudaMemcpy(h, d, sizeof(h[0]) * 96, cudaMemcpyDeviceToHost);
  int o = 0;
  const int Ns[3] = {16, 32, 64};
  for (int ni = 0; ni < 3; ++ni) for (int mode = 0; mode < 2; ++mode) for (int nacc = 1; nacc <= 4; ++nacc) {
    printf("N=%2d A=%s accumulators=%d : issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (96 MMAs)\n", Ns[ni], mode ? "TMEM" : "SMEM", nacc, h[o] / 96.0, h[o + 1] / 96.0);
    o += 2;
  }
  cudaFuncSetAttribute(probe_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe_commit<<<1, 128, 200 * 1024>>>(d);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(h[0]) * 12, cudaMemcpyDeviceToHost);
  for (int rep = 0; rep < 2; ++rep)
    printf("48 MMAs M=128 N=128: one commit at the end: issue %lld / done %lld cyc | commit every 12 MMAs: %lld / %lld | every 6: %lld / %lld\n",
           h[rep * 6], h[rep * 6 + 1], h[rep * 6 + 2], h[rep * 6 + 3], h[rep * 6 + 4], h[rep * 6 + 5]);
  cudaFuncSetAttribute(probe_layout, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe_layout<<<1, 128, 200 * 1024>>>(d);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(h[0]) * 6, cudaMemcpyDeviceToHost);
  for (int rep = 0; rep < 2; ++rep)
    printf("48 MMAs M=128 N=128 K=16 f16 SS no-swizzle: LBO128/SBO512 %lld cyc (%.0f/MMA) | LBO160/SBO640 %lld cyc (%.0f/MMA) | LBO2048/SBO128 %lld cyc (%.0f/MMA)\n",
           h[rep * 3], h[rep * 3] / 48.0, h[rep * 3 + 1], h[rep * 3 + 1] / 48.0, h[rep * 3 + 2], h[rep * 3 + 2] / 48.0);
  probe_style<<<1, 128, 200 * 1024>>>(d);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(h[0]) * 12, cudaMemcpyDeviceToHost);
  for (int rep = 0; rep < 3; ++rep)
    printf("48 unrolled MMAs (N=16, SS): tid==0 style issue %lld complete %lld cyc | warp+elect style issue %lld complete %lld cyc\n",
           h[rep * 4], h[rep * 4 + 1], h[rep * 4 + 2], h[rep * 4 + 3]);
  {
    cudaFuncSetAttribute(probe_contend, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    uint8_t* src; cudaMalloc(&src, (size_t)8 * 148 * 16384); cudaMemset(src, 0, (size_t)8 * 148 * 16384);
    long long* dd; cudaMalloc(&dd, 148 * 2 * sizeof(long long));
    const int R = 16, C = 96;     // 768 MMAs (= 16 tiles' worth of k16 x 3 products) | 96 x 16 KB
    const int modes[9] = {1, 2, 3, 5, 7, 9, 17, 25, 27};
    const char* names[9] = {"MMA SS alone", "bulk copies alone", "MMA SS + bulk copies", "MMA TS alone", "MMA TS + bulk copies",
                            "MMA SS random data", "MMA SS + 24 spinning warps", "MMA SS random + spinning", "MMA SS random+spin+copies"};
    for (int grid : {1, 148})
      for (int mi = 0; mi < 9; ++mi) {
        long long hh[296];
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(dd, 0, sizeof(hh));
          probe_contend<<<grid, (modes[mi] & 16) ? 832 : 128, 200 * 1024>>>(dd, src, modes[mi], R, C);
          e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(hh, dd, sizeof(hh), cudaMemcpyDeviceToHost);
        double m = 0, c = 0;
        for (int b = 0; b < grid; ++b) { m += hh[2 * b]; c += hh[2 * b + 1]; }
        printf("contend grid=%3d %-27s: MMA chain %8.0f cyc (%.1f/MMA) | copies %8.0f cyc (%.0f cyc per 16 KB)\n", grid, names[mi], m / grid,
               m / grid / (48.0 * R), c / grid, c / grid / C);
      }
  }
  {
    long long hh[1];
    auto run = [&](auto kfn, const char* name) {
      cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
      for (int rep = 0; rep < 2; ++rep) kfn<<<148, 128, 216 * 1024>>>(d, 20, 8, 5);
      cudaError_t e2 = cudaDeviceSynchronize();
      if (e2 != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e2)); return; }
      cudaMemcpy(hh, d, sizeof(hh), cudaMemcpyDeviceToHost);
      printf("loop form %-58s: %8lld cyc for 960 MMAs = %.1f cyc/MMA\n", name, hh[0], hh[0] / 960.0);
    };
    run(probe_loop<0>, "V0 elect per k-block, runtime ring addresses");
    run(probe_loop<1>, "V1 one thread runs the loop, runtime ring addresses");
    run(probe_loop<3>, "V3 = V1 unrolled by 4");
    run(probe_loop<2>, "V2 descriptors precomputed, 40 k-blocks unrolled");
  }
  {
    cudaFuncSetAttribute(probe_two_issuers, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe_two_issuers<<<1, 128, 200 * 1024>>>(d);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h[0]) * 6, cudaMemcpyDeviceToHost);
    for (int rep = 0; rep < 3; ++rep)
      printf("two issuers: 48 MMAs (N=16) from one thread %lld cyc | split over two warps %lld cyc\n", h[rep * 2], h[rep * 2 + 1]);
  }
  return 0;
}
