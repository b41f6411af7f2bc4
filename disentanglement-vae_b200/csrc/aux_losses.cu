// Auxiliary disentanglement objectives on the latent spaces (see include/dvae_b200.h):
//   dvae_entropy_loss   -- AdversarialDiscriminator.compute_adversarial_loss (vae/model.py:247-258)
//   dvae_club_mi        -- CLUB.forward, the MI upper bound (vae/losses.py:53-67)
//   dvae_club_nll       -- CLUB.learning_loss = -loglikeli (vae/losses.py:69-74)
//   dvae_act_bwd        -- tanh / ReLU derivative for the estimators' two-layer MLPs
// All operands are [B, <= 64-ish] matrices: a few KB.  One CTA, fixed-order reductions (bit-reproducible), forward
// and backward in the same entry point (backward when the gradient outputs are given).  Launch-latency bound.
#include "common.cuh"

namespace dvae {
namespace {

constexpr int kAuxThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < kAuxThreads / 32; ++i) t += red[i];
  return t;
}

// loss = mean_b sum_c p log p,  p = clamp(sigmoid(x) | softmax(x), 1e-8, 1 - 1e-8)      (= -mean entropy term)
__global__ void __launch_bounds__(kAuxThreads) entropy_loss_kernel(const float* __restrict__ logits, int B, int O,
                                                                   float* __restrict__ loss, const float* __restrict__ g_loss,
                                                                   float* __restrict__ d_logits) {
  __shared__ float red[kAuxThreads / 32];
  const float lo = 1e-8f, hi = 1.f - 1e-8f;
  const float g = d_logits ? (g_loss ? g_loss[0] : 1.f) / (float)B : 0.f;
  float local = 0.f;
  for (int b = threadIdx.x; b < B; b += kAuxThreads) {
    const float* x = logits + (int64_t)b * O;
    if (O == 1) {
      const float p = sigmoidf_(x[0]);
      const float pc = fminf(fmaxf(p, lo), hi);
      local += pc * logf(pc);
      if (d_logits) d_logits[b] = (p >= lo && p <= hi) ? g * (logf(pc) + 1.f) * p * (1.f - p) : 0.f;
    } else {
      float mx = x[0];
      for (int c = 1; c < O; ++c) mx = fmaxf(mx, x[c]);
      float se = 0.f;
      for (int c = 0; c < O; ++c) se += expf(x[c] - mx);
      float row = 0.f, dot = 0.f;      // dot = sum_c p_c * dL/dp_c
      for (int c = 0; c < O; ++c) {
        const float p = expf(x[c] - mx) / se;
        const float pc = fminf(fmaxf(p, lo), hi);
        row += pc * logf(pc);
        if (p >= lo && p <= hi) dot += p * (logf(pc) + 1.f);
      }
      local += row;
      if (d_logits)
        for (int c = 0; c < O; ++c) {
          const float p = expf(x[c] - mx) / se;
          const float pc = fminf(fmaxf(p, lo), hi);
          const float dldp = (p >= lo && p <= hi) ? logf(pc) + 1.f : 0.f;
          d_logits[(int64_t)b * O + c] = g * p * (dldp - dot);
        }
    }
  }
  const float tot = block_sum(local, red);
  if (threadIdx.x == 0 && loss) loss[0] = tot / (float)B;
}

// CLUB upper bound: mi = mean_i sum_d [ -(mu_id - y_id)^2 + mean_j (y_jd - mu_id)^2 ] / (2 exp(logvar_id))
__global__ void __launch_bounds__(kAuxThreads) club_mi_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                                              const float* __restrict__ y, int B, int D, float* __restrict__ mi,
                                                              const float* __restrict__ g_mi, float* __restrict__ d_mu,
                                                              float* __restrict__ d_logvar, float* __restrict__ d_y,
                                                              float* __restrict__ colstat /* [3][D] scratch */) {
  __shared__ float red[kAuxThreads / 32];
  const int n = B * D;
  const float invB = 1.f / (float)B;
  float local = 0.f;
  for (int i = threadIdx.x; i < n; i += kAuxThreads) {
    const int d = i % D;
    const float m = mu[i], w = expf(-logvar[i]);
    float msq = 0.f;                                   // mean_j (y_jd - mu_id)^2, summed in row order like the reference
    for (int j = 0; j < B; ++j) {
      const float t = y[(int64_t)j * D + d] - m;
      msq = fmaf(t, t, msq);
    }
    msq *= invB;
    const float diff = m - y[i];
    local += 0.5f * w * (msq - diff * diff);
  }
  const float tot = block_sum(local, red);
  if (threadIdx.x == 0 && mi) mi[0] = tot * invB;
  if (!d_mu) return;
  // backward.  column statistics: Ey_d, W_d = mean_i exp(-lv_id), M_d = mean_i mu_id exp(-lv_id)
  for (int d = threadIdx.x; d < D; d += kAuxThreads) {
    float ey = 0.f, wsum = 0.f, msum = 0.f;
    for (int j = 0; j < B; ++j) {
      const float w = expf(-logvar[(int64_t)j * D + d]);
      ey += y[(int64_t)j * D + d];
      wsum += w;
      msum = fmaf(mu[(int64_t)j * D + d], w, msum);
    }
    colstat[d] = ey * invB; colstat[D + d] = wsum * invB; colstat[2 * D + d] = msum * invB;
  }
  __syncthreads();
  const float g = (g_mi ? g_mi[0] : 1.f) * invB;
  for (int i = threadIdx.x; i < n; i += kAuxThreads) {
    const int d = i % D;
    const float m = mu[i], w = expf(-logvar[i]), yy = y[i];
    float msq = 0.f;
    for (int j = 0; j < B; ++j) {
      const float t = y[(int64_t)j * D + d] - m;
      msq = fmaf(t, t, msq);
    }
    msq *= invB;
    const float diff = m - yy;
    d_mu[i] = g * (yy - colstat[d]) * w;                               // d/dmu [(-(mu-y)^2 + msq) / (2 e^lv)]
    d_logvar[i] = -g * 0.5f * w * (msq - diff * diff);                 // d/dlv = -(term)
    d_y[i] = g * (diff * w + yy * colstat[D + d] - colstat[2 * D + d]);
  }
}

// CLUB learning loss: nll = -mean_b sum_d [ -(mu - y)^2 / exp(logvar) - logvar ]
__global__ void __launch_bounds__(kAuxThreads) club_nll_kernel(const float* __restrict__ mu, const float* __restrict__ logvar,
                                                               const float* __restrict__ y, int B, int D, float* __restrict__ loss,
                                                               const float* __restrict__ g_loss, float* __restrict__ d_mu,
                                                               float* __restrict__ d_logvar) {
  __shared__ float red[kAuxThreads / 32];
  const int n = B * D;
  const float invB = 1.f / (float)B;
  const float g = d_mu ? (g_loss ? g_loss[0] : 1.f) * invB : 0.f;
  float local = 0.f;
  for (int i = threadIdx.x; i < n; i += kAuxThreads) {
    const float diff = mu[i] - y[i], lv = logvar[i], w = expf(-lv);
    local += diff * diff * w + lv;
    if (d_mu) {
      d_mu[i] = g * 2.f * diff * w;
      d_logvar[i] = g * (1.f - diff * diff * w);
    }
  }
  const float tot = block_sum(local, red);
  if (threadIdx.x == 0 && loss) loss[0] = tot * invB;
}

__global__ void act_bwd_kernel(const float* __restrict__ yv, const float* __restrict__ g, float* __restrict__ d, int64_t n, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = yv[i];
  d[i] = act == 1 ? g[i] * (1.f - v * v) : (v > 0.f ? g[i] : 0.f);
}

__global__ void relu_kernel(float* __restrict__ x, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmaxf(x[i], 0.f);
}

}  // namespace
}  // namespace dvae

extern "C" int dvae_entropy_loss(const float* logits, int B, int O, float* loss, const float* g_loss, float* d_logits, void* stream) {
  DVAE_REQUIRE(logits && B > 0 && O > 0 && (loss || d_logits), "dvae_entropy_loss: bad argument");
  dvae::entropy_loss_kernel<<<1, dvae::kAuxThreads, 0, (cudaStream_t)stream>>>(logits, B, O, loss, g_loss, d_logits);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

extern "C" int dvae_club_mi(const float* mu, const float* logvar, const float* y, int B, int D, float* mi, const float* g_mi,
                            float* d_mu, float* d_logvar, float* d_y, float* ws, void* stream) {
  DVAE_REQUIRE(mu && logvar && y && B > 0 && D > 0 && (mi || d_mu), "dvae_club_mi: bad argument");
  DVAE_REQUIRE(!d_mu || (d_logvar && d_y && ws), "dvae_club_mi: backward needs d_mu, d_logvar, d_y and 3*D floats of workspace");
  dvae::club_mi_kernel<<<1, dvae::kAuxThreads, 0, (cudaStream_t)stream>>>(mu, logvar, y, B, D, mi, g_mi, d_mu, d_logvar, d_y, ws);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

extern "C" int dvae_club_nll(const float* mu, const float* logvar, const float* y, int B, int D, float* loss, const float* g_loss,
                             float* d_mu, float* d_logvar, void* stream) {
  DVAE_REQUIRE(mu && logvar && y && B > 0 && D > 0 && (loss || d_mu), "dvae_club_nll: bad argument");
  DVAE_REQUIRE(!d_mu || d_logvar, "dvae_club_nll: backward needs d_mu and d_logvar");
  dvae::club_nll_kernel<<<1, dvae::kAuxThreads, 0, (cudaStream_t)stream>>>(mu, logvar, y, B, D, loss, g_loss, d_mu, d_logvar);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

extern "C" int dvae_act_bwd(const float* y, const float* g, float* d, int64_t n, int act, void* stream) {
  DVAE_REQUIRE(y && g && d && n > 0 && (act == 1 || act == 2), "dvae_act_bwd: bad argument");
  dvae::act_bwd_kernel<<<dvae::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(y, g, d, n, act);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}

extern "C" int dvae_relu(float* x, int64_t n, void* stream) {
  DVAE_REQUIRE(x && n > 0, "dvae_relu: bad argument");
  dvae::relu_kernel<<<dvae::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n);
  DVAE_LAUNCH_CHECK();
  return DVAE_OK;
}
