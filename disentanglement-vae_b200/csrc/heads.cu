// Fused latent heads (see include/dvae_b200.h, dvae_latent_heads_fwd/bwd).
//
// Forward is ONE kernel: context -> (mu, raw) projection for every latent space, logvar = tanh(raw),
// reparameterised sample z = mu + eps * exp(logvar), per-space KL and its lambda / cyclic weight,
// discriminator logits + BCE/CE loss + accuracy, and the z -> decoder-initial-state projection
// tanh(z2hidden(z)).  Each CTA owns kRows batch rows (context rows staged in shared memory, weight
// rows streamed coalesced from L2, warp-shuffle reductions); per-CTA partial sums are combined in a
// fixed order by the last CTA to finish, so the scalars are bit-reproducible.
#include <stdlib.h>
#include "common.cuh"

namespace dvae {

int linear_impl(const float* A, int64_t lda, int trans_a, const float* B, int64_t ldb, int trans_b, float* C,
                int64_t ldc, int M, int N, int K, const float* bias, const float* bias2, float beta, int act,
                cudaStream_t st);
int colsum_impl(const float* X, int64_t ldx, int M, int N, float* out, float beta, cudaStream_t st);

constexpr int kRows = 2;
constexpr int kHeadsThreads = 1024;   // 32 warps: all 128 (mu, raw) columns in one pass, one z2hidden column per thread -- every CTA streams
                                      // the full weight matrices from L2, so the number of loads in flight per CTA is what sets its time
constexpr int kMaxDscOut = 64;

struct HeadsMeta {
  int S, Z, OD;                       // spaces, total latent dim, total discriminator outputs
  int zdim[DVAE_MAX_SPACES];
  int zoff[DVAE_MAX_SPACES];          // column offset of the space inside [B,Z]
  int dout[DVAE_MAX_SPACES];          // discriminator output dim (0 = none)
  int doff[DVAE_MAX_SPACES];          // column offset inside dsc_logits / b_dsc
  int dwoff[DVAE_MAX_SPACES];         // float offset inside w_dsc
  int dlab[DVAE_MAX_SPACES];          // index of the label row ([n_dsc][B])
};

static int make_meta(int S, const int* space_dims, const int* dsc_out, HeadsMeta* m) {
  DVAE_REQUIRE(S > 0 && S <= DVAE_MAX_SPACES, "latent heads: S=%d out of range", S);
  m->S = S;
  int z = 0, od = 0, w = 0, nl = 0;
  for (int s = 0; s < S; ++s) {
    DVAE_REQUIRE(space_dims[s] > 0, "latent heads: space %d has dim %d", s, space_dims[s]);
    m->zdim[s] = space_dims[s]; m->zoff[s] = z; z += space_dims[s];
    int o = dsc_out ? dsc_out[s] : 0;
    m->dout[s] = o; m->doff[s] = od; m->dwoff[s] = w; m->dlab[s] = o > 0 ? nl : -1;
    if (o > 0) { od += o; w += o * space_dims[s]; ++nl; }
  }
  DVAE_REQUIRE(od <= kMaxDscOut, "latent heads: %d discriminator outputs > %d", od, kMaxDscOut);
  m->Z = z; m->OD = od;
  return DVAE_OK;
}

struct HeadsFwdArgs {
  const float *ctx, *w_c2p, *b_c2p, *eps, *w_dsc, *b_dsc, *labels, *kl_w, *w_z2h, *b_z2h;
  float *z, *mu, *logvar, *hid, *dsc_logits, *scalars, *ws;
  int B, C, H2L;
  unsigned long long* dbg;       // probes (DVAE_HEADS_DBG=<device address>): globaltimer marks of CTA 0
};
#define HEADS_MARK(i) do { if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.dbg[i] = t_; } } while (0)

__device__ __forceinline__ int space_of(const HeadsMeta& m, int zi) {
  int s = 0;
  while (s + 1 < m.S && zi >= m.zoff[s + 1]) ++s;
  return s;
}

// ROWS batch rows per CTA: 2 at training batch sizes (more CTAs, shorter critical path); 4 for the large inference
// batches, where the per-CTA pass over the full weight matrices, not the row count, sets the time (B = 1024: 512 CTAs x 768 KB
// from L2 took 193 us)
template <int ROWS>
__global__ void __launch_bounds__(kHeadsThreads) heads_fwd_kernel(HeadsFwdArgs a, HeadsMeta m) {
  extern __shared__ __align__(16) float sm[];
  const int C = a.C, Z = m.Z, S = m.S, OD = m.OD, B = a.B;
  float* ctx_s = sm;                         // [ROWS][C]
  float* par_s = ctx_s + ROWS * C;          // [ROWS][2Z]
  float* z_s = par_s + ROWS * 2 * Z;        // [ROWS][Z]
  float* klt_s = z_s + ROWS * Z;            // [ROWS][Z] KL terms
  float* lg_s = klt_s + ROWS * Z;           // [ROWS][OD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kHeadsThreads / 32;
  const int b0 = blockIdx.x * ROWS;
  HEADS_MARK(0);

  for (int i = tid; i < ROWS * C; i += kHeadsThreads) {
    int r = i / C, k = i % C;
    ctx_s[i] = (b0 + r < B) ? a.ctx[(int64_t)(b0 + r) * C + k] : 0.f;
  }
  __syncthreads();
  HEADS_MARK(1);
  // (mu, raw) projections: each warp owns 4 output columns per pass (4 independent coalesced weight streams in
  // flight per lane), lanes stride the context width, shuffle reduction
  for (int j0 = warp * 4; j0 < 2 * Z; j0 += nwarp * 4) {
    const int nj = min(4, 2 * Z - j0);
    const float* w = a.w_c2p + (int64_t)j0 * C;
    float acc[4][ROWS] = {};
    if ((C & 3) == 0 && (reinterpret_cast<uintptr_t>(a.w_c2p) & 15) == 0) {
      // 16-byte loads: 4 weight rows x 4 k per lane and iteration, two iterations in flight
#pragma unroll 2
      for (int k = lane * 4; k < C; k += 128) {
        float4 wv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          wv[q] = q < nj ? __ldg(reinterpret_cast<const float4*>(w + (int64_t)q * C + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const float4 c = *reinterpret_cast<const float4*>(ctx_s + r * C + k);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[q][r] = fmaf(c.x, wv[q].x, acc[q][r]); acc[q][r] = fmaf(c.y, wv[q].y, acc[q][r]);
            acc[q][r] = fmaf(c.z, wv[q].z, acc[q][r]); acc[q][r] = fmaf(c.w, wv[q].w, acc[q][r]);
          }
        }
      }
    } else {
      for (int k = lane; k < C; k += 32) {
        float wv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wv[q] = q < nj ? w[(int64_t)q * C + k] : 0.f;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const float c = ctx_s[r * C + k];
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q][r] = fmaf(c, wv[q], acc[q][r]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        float v = warp_sum(acc[q][r]);
        if (lane == 0 && q < nj) par_s[r * 2 * Z + j0 + q] = v + a.b_c2p[j0 + q];
      }
  }
  __syncthreads();
  HEADS_MARK(2);
  // reparameterisation + KL terms
  for (int i = tid; i < ROWS * Z; i += kHeadsThreads) {
    int r = i / Z, zi = i % Z, b = b0 + r;
    int s = space_of(m, zi), d = zi - m.zoff[s];
    float mu = par_s[r * 2 * Z + 2 * m.zoff[s] + d];
    float lv = tanhf(par_s[r * 2 * Z + 2 * m.zoff[s] + m.zdim[s] + d]);
    float e = expf(lv);
    float zz = 0.f, kt = 0.f;
    if (b < B) {
      zz = fmaf(a.eps[(int64_t)b * Z + zi], e, mu);
      kt = 0.5f * (e + mu * mu - 1.f - lv);
      a.z[(int64_t)b * Z + zi] = zz;
      a.mu[(int64_t)b * Z + zi] = mu;
      a.logvar[(int64_t)b * Z + zi] = lv;
    }
    z_s[i] = zz;
    klt_s[i] = kt;
  }
  __syncthreads();
  // discriminator logits
  for (int i = tid; i < ROWS * OD; i += kHeadsThreads) {
    int r = i / OD, od = i % OD, b = b0 + r;
    int s = 0;
    while (!(m.dout[s] > 0 && od >= m.doff[s] && od < m.doff[s] + m.dout[s])) ++s;
    int o = od - m.doff[s];
    const float* w = a.w_dsc + m.dwoff[s] + o * m.zdim[s];
    float acc = a.b_dsc[od];
    for (int d = 0; d < m.zdim[s]; ++d) acc = fmaf(z_s[r * Z + m.zoff[s] + d], w[d], acc);
    lg_s[i] = acc;
    if (b < B) a.dsc_logits[(int64_t)b * OD + od] = acc;
  }
  HEADS_MARK(3);
  // decoder initial state: hid = tanh(z . Wz^T + bz); one thread per output column, vectorised weight row
  for (int j = tid; j < a.H2L; j += kHeadsThreads) {
    const float* w = a.w_z2h + (int64_t)j * Z;
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = a.b_z2h[j];
    if ((Z & 3) == 0 && (reinterpret_cast<uintptr_t>(a.w_z2h) & 15) == 0) {
      for (int k = 0; k < Z; k += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(w + k);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          const float* zr = z_s + r * Z + k;
          acc[r] = fmaf(zr[0], wv.x, acc[r]); acc[r] = fmaf(zr[1], wv.y, acc[r]);
          acc[r] = fmaf(zr[2], wv.z, acc[r]); acc[r] = fmaf(zr[3], wv.w, acc[r]);
        }
      }
    } else {
      for (int k = 0; k < Z; ++k) {
        const float wv = w[k];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(z_s[r * Z + k], wv, acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      if (b0 + r < B) a.hid[(int64_t)(b0 + r) * a.H2L + j] = tanhf(acc[r]);
  }
  __syncthreads();
  HEADS_MARK(4);
  // per-CTA partial sums, fixed order: thread s handles space s
  if (tid < S) {
    const int s = tid;
    float kl = 0.f, dl = 0.f, da = 0.f;
    for (int r = 0; r < ROWS; ++r) {
      if (b0 + r >= B) break;
      for (int d = 0; d < m.zdim[s]; ++d) kl += klt_s[r * Z + m.zoff[s] + d];
      if (m.dout[s] > 0 && a.labels) {
        const float y = a.labels[(int64_t)m.dlab[s] * B + b0 + r];
        const float* l = lg_s + r * OD + m.doff[s];
        if (m.dout[s] == 1) {
          float x = l[0];
          dl += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
          da += ((x > 0.f ? 1.f : 0.f) == y) ? 1.f : 0.f;
        } else {
          float mx = l[0];
          int am = 0;
          for (int o = 1; o < m.dout[s]; ++o)
            if (l[o] > mx) { mx = l[o]; am = o; }
          float se = 0.f;
          for (int o = 0; o < m.dout[s]; ++o) se += expf(l[o] - mx);
          int yi = (int)y;
          dl += mx + logf(se) - l[yi];
          da += (am == yi) ? 1.f : 0.f;
        }
      }
    }
    float* part = a.ws + (int64_t)blockIdx.x * 3 * S;
    part[s] = kl; part[S + s] = dl; part[2 * S + s] = da;
  }
  HEADS_MARK(5);
  // last CTA combines the partials in block order
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    unsigned int* counter = reinterpret_cast<unsigned int*>(a.ws + (int64_t)gridDim.x * 3 * S);
    // atomicInc wraps to 0 on the last arrival: the counter re-arms itself (no memset node in front of the kernel);
    // the caller zero-fills ws once before the first call
    is_last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
  }
  __syncthreads();
  // all threads of the last CTA fetch the partials into shared memory in parallel; thread 0 then adds them in block
  // order from there (a single thread walking 3*S*grid dependent L2 loads was a serial tail of this kernel)
  const int n_part = (int)gridDim.x * 3 * S;
  const bool staged = n_part <= ROWS * C;          // ctx_s is free by now
  if (is_last && staged) {
    __threadfence();
    for (int i = tid; i < n_part; i += kHeadsThreads) ctx_s[i] = __ldcg(a.ws + i);
  }
  __syncthreads();
  if (is_last && tid == 0) {
    __threadfence();
    float wkl = 0.f, tkl = 0.f, tdl = 0.f;
    for (int s = 0; s < S; ++s) {
      float kl = 0.f, dl = 0.f, da = 0.f;
      for (unsigned int c = 0; c < gridDim.x; ++c) {
        if (staged) {
          const float* part = ctx_s + c * 3 * S;
          kl += part[s]; dl += part[S + s]; da += part[2 * S + s];
        } else {
          const volatile float* part = a.ws + (int64_t)c * 3 * S;
          kl += part[s]; dl += part[S + s]; da += part[2 * S + s];
        }
      }
      kl /= (float)B;
      dl /= (float)(B * (m.dout[s] == 1 ? 1 : 1));
      da /= (float)B;
      a.scalars[3 + s] = kl;
      a.scalars[3 + S + s] = m.dout[s] > 0 ? dl : 0.f;
      a.scalars[3 + 2 * S + s] = m.dout[s] > 0 ? da : 0.f;
      wkl = fmaf(a.kl_w ? a.kl_w[s] : 0.f, kl, wkl);
      tkl += kl;
      if (m.dout[s] > 0) tdl += dl;
    }
    a.scalars[0] = wkl; a.scalars[1] = tkl; a.scalars[2] = tdl;
  }
  HEADS_MARK(6);
}

// ---- backward ---------------------------------------------------------------------------------
__global__ void heads_bwd_pre_kernel(const float* __restrict__ d_hid, const float* __restrict__ hid,
                                     float* __restrict__ d_pre, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float h = hid[i]; d_pre[i] = d_hid[i] * (1.f - h * h); }
}

struct HeadsBwdArgs {
  const float *eps, *w_dsc, *labels, *kl_w, *z, *mu, *logvar, *dsc_logits, *d_z, *d_z_extra;
  const float *d_mu_extra, *d_logvar_extra, *d_logits_extra;
  float *dp, *dl;     // [B,2Z], [B,OD]
  int B;
};

// one thread per (b, zi); the first thread of each (b, space-with-dsc) also emits dl[b][*]
__global__ void heads_bwd_elem_kernel(HeadsBwdArgs a, HeadsMeta m) {
  const int Z = m.Z, OD = m.OD, B = a.B;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * Z) return;
  const int b = i / Z, zi = i % Z;
  const int s = space_of(m, zi), d = zi - m.zoff[s], zs = m.zdim[s];
  float dz = a.d_z[i] + (a.d_z_extra ? a.d_z_extra[i] : 0.f);
  if (m.dout[s] > 0) {
    const float* w = a.w_dsc + m.dwoff[s];
    const float* l = a.dsc_logits + (int64_t)b * OD + m.doff[s];
    const float* gx = a.d_logits_extra ? a.d_logits_extra + (int64_t)b * OD + m.doff[s] : nullptr;
    float mx = 0.f, se = 1.f, y = 0.f;
    if (a.labels) {
      y = a.labels[(int64_t)m.dlab[s] * B + b];
      if (m.dout[s] > 1) {
        mx = l[0];
        for (int o = 1; o < m.dout[s]; ++o) mx = fmaxf(mx, l[o]);
        se = 0.f;
        for (int o = 0; o < m.dout[s]; ++o) se += expf(l[o] - mx);
      }
    }
    for (int o = 0; o < m.dout[s]; ++o) {
      float g = gx ? gx[o] : 0.f;
      if (a.labels) {
        if (m.dout[s] == 1) g += (sigmoidf_(l[0]) - y) / (float)B;
        else g += (expf(l[o] - mx) / se - (o == (int)y ? 1.f : 0.f)) / (float)B;
      }
      dz = fmaf(g, w[o * zs + d], dz);
      if (d == 0) a.dl[(int64_t)b * OD + m.doff[s] + o] = g;
    }
  }
  const float w_kl = a.kl_w ? a.kl_w[s] : 0.f;
  const float mu = a.mu[i], lv = a.logvar[i], e = expf(lv);
  const float dmu = fmaf(w_kl / (float)B, mu, dz) + (a.d_mu_extra ? a.d_mu_extra[i] : 0.f);
  const float dlv = fmaf(dz * a.eps[i], e, w_kl * 0.5f * (e - 1.f) / (float)B) + (a.d_logvar_extra ? a.d_logvar_extra[i] : 0.f);
  a.dp[(int64_t)b * 2 * Z + 2 * m.zoff[s] + d] = dmu;
  a.dp[(int64_t)b * 2 * Z + 2 * m.zoff[s] + zs + d] = dlv * (1.f - lv * lv);
}

// d_w_dsc[o][d] = sum_b dl[b][o] z[b][zoff+d]; d_b_dsc[o] = sum_b dl[b][o].  One warp per element.
__global__ void heads_bwd_dsc_kernel(const float* __restrict__ dl, const float* __restrict__ z, float* d_w_dsc,
                                     float* d_b_dsc, int B, HeadsMeta m, int n_w) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_w + m.OD) return;
  float acc = 0.f;
  if (warp < n_w) {
    int s = 0;
    while (!(m.dout[s] > 0 && warp >= m.dwoff[s] && warp < m.dwoff[s] + m.dout[s] * m.zdim[s])) ++s;
    int o = (warp - m.dwoff[s]) / m.zdim[s], d = (warp - m.dwoff[s]) % m.zdim[s];
    for (int b = lane; b < B; b += 32) acc = fmaf(dl[(int64_t)b * m.OD + m.doff[s] + o], z[(int64_t)b * m.Z + m.zoff[s] + d], acc);
    acc = warp_sum(acc);
    if (lane == 0) d_w_dsc[warp] = acc;
  } else {
    int od = warp - n_w;
    for (int b = lane; b < B; b += 32) acc += dl[(int64_t)b * m.OD + od];
    acc = warp_sum(acc);
    if (lane == 0) d_b_dsc[od] = acc;
  }
}

// Standalone discriminator loss (used when labels are not handed to the fused forward):
// out[s] = loss, out[S+s] = accuracy; d_logits optional.
__global__ void dsc_loss_kernel(const float* __restrict__ logits, const float* __restrict__ labels, int B, HeadsMeta m,
                                float* __restrict__ out, const float* __restrict__ d_out, float* __restrict__ d_logits) {
  const int s = blockIdx.x, OD = m.OD;
  if (m.dout[s] == 0) { if (out && threadIdx.x == 0) { out[s] = 0.f; out[m.S + s] = 0.f; } return; }
  __shared__ float red[2][8];
  float ls = 0.f, ac = 0.f;
  const float go = d_out ? d_out[s] / (float)B : 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* l = logits + (int64_t)b * OD + m.doff[s];
    const float y = labels[(int64_t)m.dlab[s] * B + b];
    if (m.dout[s] == 1) {
      float x = l[0];
      ls += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
      ac += ((x > 0.f ? 1.f : 0.f) == y) ? 1.f : 0.f;
      if (d_logits) d_logits[(int64_t)b * OD + m.doff[s]] = (sigmoidf_(x) - y) * go;
    } else {
      float mx = l[0];
      int am = 0;
      for (int o = 1; o < m.dout[s]; ++o)
        if (l[o] > mx) { mx = l[o]; am = o; }
      float se = 0.f;
      for (int o = 0; o < m.dout[s]; ++o) se += expf(l[o] - mx);
      const int yi = (int)y;
      ls += mx + logf(se) - l[yi];
      ac += (am == yi) ? 1.f : 0.f;
      if (d_logits)
        for (int o = 0; o < m.dout[s]; ++o) d_logits[(int64_t)b * OD + m.doff[s] + o] = (expf(l[o] - mx) / se - (o == yi ? 1.f : 0.f)) * go;
    }
  }
  ls = warp_sum(ls); ac = warp_sum(ac);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ls; red[1][threadIdx.x >> 5] = ac; }
  __syncthreads();
  if (threadIdx.x == 0 && out) {
    float a0 = 0.f, a1 = 0.f;
    for (int i = 0; i < 8; ++i) { a0 += red[0][i]; a1 += red[1][i]; }
    out[s] = a0 / (float)B; out[m.S + s] = a1 / (float)B;
  }
}

}  // namespace dvae

using namespace dvae;

extern "C" int dvae_dsc_loss(const float* dsc_logits, const float* labels, int B, int S, const int* space_dims,
                             const int* dsc_out, float* out, const float* d_out, float* d_logits, void* stream) {
  DVAE_REQUIRE(dsc_logits && labels && B > 0 && (out || d_logits), "dvae_dsc_loss: bad argument");
  HeadsMeta m;
  int rc = make_meta(S, space_dims, dsc_out, &m);
  if (rc) return rc;
  dsc_loss_kernel<<<S, 256, 0, (cudaStream_t)stream>>>(dsc_logits, labels, B, m, out, d_out, d_logits);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

extern "C" int64_t dvae_heads_ws_floats(int B, int S) { return (int64_t)ceil_div(B, kRows) * 3 * S + 4; }

extern "C" int dvae_latent_heads_fwd(const float* ctx, int B, int C, int S, const int* space_dims,
                                     const int* dsc_out, const float* w_c2p, const float* b_c2p, const float* eps,
                                     const float* w_dsc, const float* b_dsc, const float* labels,
                                     const float* kl_w_dev, const float* w_z2h, const float* b_z2h, int H2L,
                                     float* z, float* mu, float* logvar, float* hid, float* dsc_logits,
                                     float* scalars, float* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DVAE_REQUIRE(ctx && space_dims && w_c2p && b_c2p && eps && w_z2h && b_z2h && z && mu && logvar && hid && scalars && ws,
               "dvae_latent_heads_fwd: null pointer");
  DVAE_REQUIRE(B > 0 && C > 0 && H2L > 0, "dvae_latent_heads_fwd: bad shape");
  HeadsMeta m;
  int rc = make_meta(S, space_dims, dsc_out, &m);
  if (rc) return rc;
  DVAE_REQUIRE(m.OD == 0 || (w_dsc && b_dsc && dsc_logits), "dvae_latent_heads_fwd: discriminator buffers missing");
  HeadsFwdArgs a{ctx, w_c2p, b_c2p, eps, w_dsc, b_dsc, labels, kl_w_dev, w_z2h, b_z2h,
                 z, mu, logvar, hid, dsc_logits, scalars, ws, B, C, H2L, nullptr};
  if (const char* e = getenv("DVAE_HEADS_DBG")) a.dbg = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
  auto smem_for = [&](int rows) { return sizeof(float) * ((size_t)rows * C + rows * 4 * m.Z + rows * (m.OD > 0 ? m.OD : 1) + 3 * S); };
  int rows = B > 256 ? 4 : kRows;      // 8 rows per CTA spills at 1024 threads (64 registers each)
  while (rows > kRows && smem_for(rows) > 160 * 1024) rows /= 2;
  const size_t smem = smem_for(rows);
  DVAE_REQUIRE(smem <= 200 * 1024, "dvae_latent_heads_fwd: context width %d too large for shared memory", C);
  const int grid = ceil_div(B, rows);
#define DVAE_HEADS_LAUNCH(R)                                                                                                  \
  do {                                                                                                                        \
    if (smem > 48 * 1024) DVAE_CUDA(cudaFuncSetAttribute(heads_fwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    heads_fwd_kernel<R><<<grid, kHeadsThreads, smem, st>>>(a, m);                                                             \
  } while (0)
  if (rows == 4) DVAE_HEADS_LAUNCH(4);
  else DVAE_HEADS_LAUNCH(2);
#undef DVAE_HEADS_LAUNCH
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

extern "C" int64_t dvae_heads_bwd_ws_floats(int B, int Z, int H2L) { return (int64_t)B * (H2L + 3 * Z + kMaxDscOut); }

extern "C" int dvae_latent_heads_bwd(const float* ctx, int B, int C, int S, const int* space_dims,
                                     const int* dsc_out, const float* w_c2p, const float* eps, const float* w_dsc,
                                     const float* labels, const float* kl_w_dev, const float* w_z2h, int H2L,
                                     const float* z, const float* mu, const float* logvar, const float* hid,
                                     const float* dsc_logits, const float* d_hid, const float* d_z_extra,
                                     const float* d_mu_extra, const float* d_logvar_extra,
                                     const float* d_logits_extra, float* d_w_c2p, float* d_b_c2p, float* d_w_dsc, float* d_b_dsc,
                                     float* d_w_z2h, float* d_b_z2h, float* d_ctx, float* ws, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DVAE_REQUIRE(ctx && space_dims && w_c2p && eps && w_z2h && z && mu && logvar && hid && d_hid && d_w_c2p && d_b_c2p &&
                   d_w_z2h && d_b_z2h && d_ctx && ws, "dvae_latent_heads_bwd: null pointer");
  HeadsMeta m;
  int rc = make_meta(S, space_dims, dsc_out, &m);
  if (rc) return rc;
  const int Z = m.Z;
  float* d_pre = ws;                              // [B,H2L]
  float* d_z = d_pre + (int64_t)B * H2L;          // [B,Z]
  float* dp = d_z + (int64_t)B * Z;               // [B,2Z]
  float* dl = dp + (int64_t)B * 2 * Z;            // [B,OD]
  GemmHints grad_hints;            // A operands below are gradients of unknown magnitude
  grad_hints.a_wide = true;
  const int64_t n = (int64_t)B * H2L;
  heads_bwd_pre_kernel<<<ceil_div(n, 256), 256, 0, st>>>(d_hid, hid, d_pre, n);
  DVAE_LAUNCH_CHECK();
  // z2hidden: d_w = d_pre^T z, d_b = colsum(d_pre), d_z = d_pre Wz
  Fork fork(st);          // weight / bias gradients branch off the d_z -> dp -> d_ctx chain
  if ((rc = linear_impl_ex(d_pre, H2L, 1, z, Z, 1, d_w_z2h, Z, H2L, Z, B, nullptr, nullptr, 0.f, 0, grad_hints, fork.side(0)))) return rc;
  if ((rc = colsum_impl(d_pre, H2L, B, H2L, d_b_z2h, 0.f, fork.side(1)))) return rc;
  if ((rc = linear_impl_ex(d_pre, H2L, 0, w_z2h, Z, 1, d_z, Z, B, Z, H2L, nullptr, nullptr, 0.f, 0, grad_hints, st))) return rc;
  HeadsBwdArgs a{eps, w_dsc, labels, kl_w_dev, z, mu, logvar, dsc_logits, d_z, d_z_extra, d_mu_extra, d_logvar_extra, d_logits_extra, dp, dl, B};
  heads_bwd_elem_kernel<<<ceil_div((int64_t)B * Z, 256), 256, 0, st>>>(a, m);
  DVAE_LAUNCH_CHECK();
  // everything below that is not on the dp -> d_ctx chain (the encoder's backward waits for d_ctx only) runs on side stream 2:
  // discriminator weight gradients, context2params weight and bias gradients.  Side streams 0 / 1 still hold the z2hidden ones.
  {
    Fork fork2(st);
    if (m.OD > 0) {
      DVAE_REQUIRE(d_w_dsc && d_b_dsc, "dvae_latent_heads_bwd: discriminator gradient buffers missing");
      int n_w = 0;
      for (int s = 0; s < S; ++s) n_w += m.dout[s] * m.zdim[s];
      heads_bwd_dsc_kernel<<<ceil_div((int64_t)(n_w + m.OD) * 32, 256), 256, 0, fork2.side(2)>>>(dl, z, d_w_dsc, d_b_dsc, B, m, n_w);
      DVAE_LAUNCH_CHECK();
    }
    // context2params: d_w = dp^T ctx, d_b = colsum(dp), d_ctx = dp W
    if ((rc = linear_impl_ex(dp, 2 * Z, 1, ctx, C, 1, d_w_c2p, C, 2 * Z, C, B, nullptr, nullptr, 0.f, 0, grad_hints, fork2.side(2)))) return rc;
    if ((rc = colsum_impl(dp, 2 * Z, B, 2 * Z, d_b_c2p, 0.f, fork2.side(2)))) return rc;
    if ((rc = linear_impl_ex(dp, 2 * Z, 0, w_c2p, C, 1, d_ctx, C, B, C, 2 * Z, nullptr, nullptr, 0.f, 0, grad_hints, st))) return rc;
    if ((rc = fork2.join_or_defer())) return rc;
  }
  if ((rc = fork.join_or_defer())) return rc;
  return DVAE_OK;
}
