// LSTM recurrence for hidden sizes the cluster-resident kernels do not take (H = 512, 1024, ...: BASELINE cfg 4).
//
// The persistent kernels (lstm_tc.cu, lstm_persist.cu) keep a [4H/8, H] slice of W_hh per CTA on chip; at H = 1024 one
// direction's W_hh is 16 MB of fp16 (hi, lo) planes and no longer fits an 8-CTA cluster.  Here the recurrent product of
// every step runs on the tensor cores as a bulk-copy-fed GEMM over operand planes that never leave the L2:
//   * W_hh is split ONCE per call into tile-blocked fp16 (hi, lo) planes (tc_planes.cuh) -- 16 MB per direction at
//     H = 1024, resident in the 126 MB L2 for the T steps that re-read it;
//   * the carried state is WRITTEN IN PLANE FORMAT by the cell kernel of the previous step, so step t's GEMM
//       gates[t] += h_{t-1} . W_hh^T          (tc16::linear_planes: no tensor maps, no converter warps, split-K)
//     reads both operands with plain 16 KB bulk copies;
//   * an element-wise cell kernel applies the gate non-linearities, updates c / h (fp32 state), writes the saved
//     gates / cell states for the backward pass and the next step's operand planes.
// Backward mirrors it: dh_{t-1} = dG_t . W_hh over planes of W_hh^T and of the gate gradients (per-row power-of-two scale,
// undone in the GEMM epilogue), followed by the cell-backward kernel.
// Two launches per step and direction (the directions of a bidirectional layer run concurrently on two streams);
// the fp32 SIMT step kernels in lstm.cu remain for shapes this path does not take (H % 32 != 0, unaligned buffers).
//
// Reference semantics: nn.LSTM inside vae/model.py:88-101 (packed, variable length) and :152-165 (decoder).
#include <cstdlib>
#include <cstring>

#include "lstm_persist.cuh"
#include "tc_gemm16.cuh"
#include "tc_planes.cuh"

namespace dvae {

// small kernels of lstm.cu, through host wrappers (kernels are not visible across translation units without -rdc)
int transpose_launch(const float* in, float* out, int R, int C, cudaStream_t st);      // out [C,R] = in [R,C]^T
int copy_rows_launch(const float* src, int64_t lds, float* dst, int64_t ldd, int B, int H, cudaStream_t st);

namespace {

constexpr int kCellThreads = 128;

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

// ---- forward ---------------------------------------------------------------------------------------------------
// state [B,H] <- h0 / c0 (or zeros) and the step-0 operand planes of h0; rows b >= B of the planes are zero
__global__ void planes_state_init_kernel(const float* __restrict__ h0, const float* __restrict__ c0, int64_t ld0, float* __restrict__ h_state,
                                         float* __restrict__ c_state, uint8_t* __restrict__ planes, int B, int RP, int H) {
  const int HC = H / 8, idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= RP * HC) return;
  const int b = idx / HC, u0 = (idx % HC) * 8;
  float hv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, cv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (b < B) {
    if (h0) load8(h0 + (int64_t)b * ld0 + u0, hv);
    if (c0) load8(c0 + (int64_t)b * ld0 + u0, cv);
    store8(h_state + (int64_t)b * H + u0, hv);
    store8(c_state + (int64_t)b * H + u0, cv);
  }
  if (planes) tc16::store_plane_chunk(planes, b, u0, H / 32, hv, 1.f);
}

struct CellFwdArgs {
  float* gates_t;            // [B,4H] pre-activations of this step (input projection + recurrent product) -> post-activation gates
  float* cs_t;               // [B,H]
  float* hs_t; int64_t ldhs; // layer output of this step (direction's column block), zeros at padding
  float* h_state; float* c_state;      // carried fp32 state [B,H], updated in place
  const int64_t* lengths;
  uint8_t* planes;           // next step's A operand: h in plane format, [RP, H]
  int t, B, RP, H;
};

// one thread = (row, 8 consecutive hidden units): four 32-byte gate segments in, one 16-byte chunk per operand plane out
__global__ void __launch_bounds__(kCellThreads) lstm_cell_fwd_kernel(CellFwdArgs p) {
  const int H = p.H, HC = H / 8, idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.RP * HC) return;
  const int b = idx / HC, u0 = (idx % HC) * 8;
  float hv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (b < p.B) {
    float* g = p.gates_t + (int64_t)b * 4 * H + u0;
    float* hsp = p.hs_t + (int64_t)b * p.ldhs + u0;
    const int64_t sb = (int64_t)b * H + u0;
    float cv[8];
    load8(p.c_state + sb, cv);
    load8(p.h_state + sb, hv);
    const bool live = p.lengths == nullptr || p.t < p.lengths[b];
    float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (live) {
      float gi[8], gf[8], gg[8], go[8];
      load8(g, gi); load8(g + H, gf); load8(g + 2 * H, gg); load8(g + 3 * H, go);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        gi[j] = sigmoidf_(gi[j]); gf[j] = sigmoidf_(gf[j]); gg[j] = tanhf(gg[j]); go[j] = sigmoidf_(go[j]);
        cv[j] = fmaf(gf[j], cv[j], gi[j] * gg[j]);
        hv[j] = go[j] * tanhf(cv[j]);
      }
      store8(g, gi); store8(g + H, gf); store8(g + 2 * H, gg); store8(g + 3 * H, go);
      store8(p.c_state + sb, cv);
      store8(p.h_state + sb, hv);
      store8(hsp, hv);
    } else {            // frozen / padded row: zero gates and output, carried state (pack_padded_sequence semantics)
      store8(g, zero); store8(g + H, zero); store8(g + 2 * H, zero); store8(g + 3 * H, zero);
      store8(hsp, zero);
    }
    store8(p.cs_t + sb, cv);
  }
  tc16::store_plane_chunk(p.planes, b, u0, H / 32, hv, 1.f);
}

// ---- backward --------------------------------------------------------------------------------------------------
__global__ void planes_bwd_init_kernel(const float* __restrict__ d_hn, const float* __restrict__ d_cn, int64_t ldn, float* __restrict__ carry,
                                       float* __restrict__ dc, float* __restrict__ dh_rec, float* __restrict__ inv_s,
                                       unsigned* __restrict__ amax, int B, int RP, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *amax = 0u;
  if (i < RP) inv_s[i] = 1.f;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H), u = (int)(i % H);
  carry[i] = d_hn ? d_hn[(int64_t)b * ldn + u] : 0.f;
  dc[i] = d_cn ? d_cn[(int64_t)b * ldn + u] : 0.f;
  dh_rec[i] = 0.f;
}

struct CellBwdArgs {
  float* gates_t;                 // [B,4H]: post-activation gates in, gate gradients dG out
  const float* cs_t;              // [B,H] cell state of this step
  const float* c_prev; int64_t ldcp;      // cell state of the previously traversed step ([B,H] slab or c0 rows), NULL = zeros
  const float* d_hs_t; int64_t lddhs;     // upstream gradient on this step's output, NULL = none
  float* carry; float* dc;        // carried gradients [B,H] (frozen rows / initial d_hn, d_cn), updated in place
  float* dh_rec;                  // [B,H] recurrent gradient from the GEMM of this step (already unscaled); read, then cleared
  const int64_t* lengths;
  uint8_t* planes; float* inv_s;  // next step's A operand: dG rows scaled by a power of two, in plane format [RP, 4H]
  unsigned* amax;                 // running max |dG| (bit pattern) of the whole call: operand scale of the dense gradient GEMMs
  int t, B, H, use_rec;
};

// one block per row: pass 1 = cell backward (dG to the gates slab, row max), pass 2 = scaled operand planes
__global__ void __launch_bounds__(kCellThreads) lstm_cell_bwd_kernel(CellBwdArgs p) {
  __shared__ unsigned s_max[kCellThreads / 32];
  const int H = p.H, HC = H / 8, b = blockIdx.x, tid = threadIdx.x;
  const int KB = 4 * H / 32;
  if (b >= p.B) {              // padding rows of the last 128-row block: zero operand
    const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = tid; c < 4 * HC; c += blockDim.x) tc16::store_plane_chunk(p.planes, b, c * 8, KB, z, 1.f);
    if (tid == 0) p.inv_s[b] = 1.f;
    return;
  }
  const bool live = p.lengths == nullptr || p.t < p.lengths[b];
  float* g = p.gates_t + (int64_t)b * 4 * H;
  unsigned mx = 0u;
  for (int c = tid; c < HC; c += blockDim.x) {
    const int u0 = c * 8;
    const int64_t sb = (int64_t)b * H + u0;
    float rec[8], car[8], dcv[8];
    const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    load8(p.carry + sb, car);
    load8(p.dc + sb, dcv);
    if (p.use_rec) {
      load8(p.dh_rec + sb, rec);
      store8(p.dh_rec + sb, zero);             // the next step's split-K GEMM accumulates into a cleared buffer
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) rec[j] = 0.f;
    }
    if (!live) {
      float cn[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cn[j] = car[j] + rec[j];
      store8(p.carry + sb, cn);
      store8(g + u0, zero); store8(g + H + u0, zero); store8(g + 2 * H + u0, zero); store8(g + 3 * H + u0, zero);
      continue;
    }
    float gi[8], gf[8], gg[8], go[8], cc[8], cp[8], dhs[8];
    load8(g + u0, gi); load8(g + H + u0, gf); load8(g + 2 * H + u0, gg); load8(g + 3 * H + u0, go);
    load8(p.cs_t + sb, cc);
    if (p.c_prev) load8(p.c_prev + (int64_t)b * p.ldcp + u0, cp);
    if (p.d_hs_t) load8(p.d_hs_t + (int64_t)b * p.lddhs + u0, dhs);
    float o0[8], o1[8], o2[8], o3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float cpv = p.c_prev ? cp[j] : 0.f;
      const float dh = car[j] + rec[j] + (p.d_hs_t ? dhs[j] : 0.f);
      const float tc = tanhf(cc[j]);
      const float dct = fmaf(dh * go[j], 1.f - tc * tc, dcv[j]);
      o0[j] = dct * gg[j] * gi[j] * (1.f - gi[j]);
      o1[j] = dct * cpv * gf[j] * (1.f - gf[j]);
      o2[j] = dct * gi[j] * (1.f - gg[j] * gg[j]);
      o3[j] = dh * tc * go[j] * (1.f - go[j]);
      dcv[j] = dct * gf[j];
      const float am = fmaxf(fmaxf(fabsf(o0[j]), fabsf(o1[j])), fmaxf(fabsf(o2[j]), fabsf(o3[j])));
      mx = max(mx, __float_as_uint(am));
    }
    store8(p.carry + sb, zero);
    store8(p.dc + sb, dcv);
    store8(g + u0, o0); store8(g + H + u0, o1); store8(g + 2 * H + u0, o2); store8(g + 3 * H + u0, o3);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((tid & 31) == 0) s_max[tid >> 5] = mx;
  __syncthreads();               // also orders this block's dG stores before the re-reads below
  mx = 0u;
#pragma unroll
  for (int w = 0; w < kCellThreads / 32; ++w) mx = max(mx, s_max[w]);
  // row scale 2^(13 - floor(log2 max |dG[row, :]|)): exact, undone by the GEMM epilogue (c_row_scale = inv_s)
  int se = 267 - (int)(mx >> 23);
  se = se < 1 ? 1 : (se > 253 ? 253 : se);
  const float sc = __uint_as_float((unsigned)se << 23);
  if (tid == 0) {
    p.inv_s[b] = __uint_as_float((unsigned)(254 - se) << 23);
    if (mx) atomicMax(p.amax, mx);
  }
  for (int c = tid; c < 4 * HC; c += blockDim.x) {
    float v[8];
    load8(g + c * 8, v);
    tc16::store_plane_chunk(p.planes, b, c * 8, KB, v, sc);
  }
}

// gradient w.r.t. the initial state: d_h0 = carried + recurrent part of the last processed step, d_c0 = carried dc
__global__ void planes_bwd_final_kernel(const float* __restrict__ carry, const float* __restrict__ dc, const float* __restrict__ dh_rec,
                                        float* __restrict__ d_h0, float* __restrict__ d_c0, int64_t ldd0, int B, int H) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * H) return;
  const int b = (int)(i / H), u = (int)(i % H);
  if (d_h0) d_h0[(int64_t)b * ldd0 + u] = carry[i] + dh_rec[i];
  if (d_c0) d_c0[(int64_t)b * ldd0 + u] = dc[i];
}

int max_splits() {
  const char* e = getenv("DVAE_PLANES_SPLITS");      // A/B knob: cap on the split-K factor of the per-step GEMMs
  const int v = e ? atoi(e) : 16;
  return v < 1 ? 1 : v;
}

struct WsLayout {
  int64_t wpl, apl, state, dh_rec, inv_s, amax, total;      // per-direction sizes (floats) and the total over D directions
};
WsLayout ws_layout(int B, int H, int D) {
  WsLayout L;
  const int64_t a = tc16::plane_floats(4 * H, H), b = tc16::plane_floats(H, 4 * H);
  L.wpl = a > b ? a : b;
  L.apl = tc16::plane_floats(B, 4 * H);
  L.state = (int64_t)B * H;                 // x4: h, c (forward) / carry, dc (backward)
  L.dh_rec = (int64_t)B * H;
  L.inv_s = (int64_t)ceil_div(B, 128) * 128;
  L.amax = 0;
  L.total = D * (L.wpl + L.apl + L.dh_rec + L.inv_s);
  return L;
}

}  // namespace

bool planes_lstm_supported(int B, int H, int D, const void* const* ptrs, int nptr, const int64_t* lds, int nld) {
  const char* e = getenv("DVAE_LSTM_IMPL");
  if (e && (!strcmp(e, "step") || !strcmp(e, "simt"))) return false;
  if (force_simt_gemm() || !tc16::enabled()) return false;
  if (H < 128 || H % 32 != 0 || B < 1 || !(D == 1 || D == 2)) return false;
  for (int i = 0; i < nptr; ++i)
    if (ptrs[i] && (reinterpret_cast<uintptr_t>(ptrs[i]) & 15)) return false;
  for (int i = 0; i < nld; ++i)
    if (lds[i] % 4) return false;
  return true;
}

int64_t planes_lstm_ws_floats(int B, int H, int D) {
  if (H < 128 || H % 32 != 0) return 0;
  return ws_layout(B, H, D).total;
}

// recurrence of one layer (all directions) over gates that already hold the input projection; `ws` = the 4*D*B*H floats of
// carried state at the front of the caller's state workspace, `pws` = the planes_lstm_ws_floats() region
int planes_lstm_fwd(int T, int B, int H, int D, const float* const* w_hh, const float* h0, const float* c0, int64_t ld0,
                    int64_t dir0, const int64_t* lengths, float* hs, int64_t ldhs, float* hn, float* cn, int64_t ldn,
                    int64_t dirn, float* gates, float* cs, float* ws, float* pws, cudaStream_t st) {
  const WsLayout L = ws_layout(B, H, D);
  const int RP = ceil_div(B, 128) * 128, HC = H / 8;
  const int64_t slab = (int64_t)T * B * 4 * H, sf = (int64_t)B * H;
  const int cell_blocks = ceil_div((int64_t)RP * HC, kCellThreads);
  const int ms = max_splits();
  Fork fork(st);           // the two directions are independent recurrences: two streams
  for (int d = 0; d < D; ++d) {
    cudaStream_t sd = d == 0 ? st : fork.side(0);
    float* wpl = pws + d * L.wpl;
    uint8_t* apl = reinterpret_cast<uint8_t*>(pws + D * L.wpl + d * L.apl);
    float* h_state = ws + (2 * d) * sf;
    float* c_state = ws + (2 * d + 1) * sf;
    int rc = tc16::split_planes(w_hh[d], H, 4 * H, H, 1.f, wpl, sd, true);
    if (rc) return rc;
    planes_state_init_kernel<<<cell_blocks, kCellThreads, 0, sd>>>(h0 ? h0 + d * dir0 : nullptr, c0 ? c0 + d * dir0 : nullptr, ld0, h_state,
                                                                   c_state, h0 ? apl : nullptr, B, RP, H);
    DVAE_LAUNCH_CHECK();
    for (int s = 0; s < T; ++s) {
      const int t = d == 0 ? s : T - 1 - s;
      float* gates_t = gates + d * slab + (int64_t)t * B * 4 * H;
      if (s > 0 || h0) {      // gates[t] += h_{t-1} . W_hh^T   (from the zero state there is nothing to add)
        rc = tc16::linear_planes(apl, wpl, gates_t, 4 * H, B, 4 * H, H, nullptr, 1.f, 0, 1.f, 1.f, nullptr, false, ms, sd);
        if (rc) return rc;
      }
      CellFwdArgs a;
      a.gates_t = gates_t; a.cs_t = cs + ((int64_t)d * T + t) * sf; a.hs_t = hs + (int64_t)t * B * ldhs + d * H; a.ldhs = ldhs;
      a.h_state = h_state; a.c_state = c_state; a.lengths = lengths; a.planes = apl; a.t = t; a.B = B; a.RP = RP; a.H = H;
      lstm_cell_fwd_kernel<<<cell_blocks, kCellThreads, 0, sd>>>(a);
      DVAE_LAUNCH_CHECK();
    }
    if (hn && (rc = copy_rows_launch(h_state, H, hn + d * dirn, ldn, B, H, sd))) return rc;
    if (cn && (rc = copy_rows_launch(c_state, H, cn + d * dirn, ldn, B, H, sd))) return rc;
  }
  return fork.join();
}

// backward recurrence: gates (post-activation) -> dG in place; returns the device slot holding max |dG| (bit pattern)
int planes_lstm_bwd(int T, int B, int H, int D, const float* const* w_hh, const float* c0, int64_t ld0, int64_t dir0,
                    const int64_t* lengths, float* gates, const float* cs, const float* d_hs, int64_t lddhs, const float* d_hn,
                    const float* d_cn, int64_t ldn, int64_t dirn, float* d_h0, float* d_c0, int64_t ldd0, int64_t dird0,
                    float* ws, float* wt, float* pws, uint32_t* amax_slots, cudaStream_t st) {
  const WsLayout L = ws_layout(B, H, D);
  const int RP = ceil_div(B, 128) * 128;
  const int64_t slab = (int64_t)T * B * 4 * H, sf = (int64_t)B * H;
  const int ms = max_splits();
  float* base_apl = pws + D * L.wpl;
  float* base_rec = base_apl + D * L.apl;
  float* base_inv = base_rec + D * L.dh_rec;
  unsigned* amax = reinterpret_cast<unsigned*>(amax_slots);      // [D] entries, owned (and rotated) by the caller
  Fork fork(st);
  for (int d = 0; d < D; ++d) {
    cudaStream_t sd = d == 0 ? st : fork.side(0);
    float* wpl = pws + d * L.wpl;
    uint8_t* apl = reinterpret_cast<uint8_t*>(base_apl + d * L.apl);
    float* dh_rec = base_rec + d * L.dh_rec;
    float* inv_s = base_inv + d * L.inv_s;
    float* carry = ws + (2 * d) * sf;
    float* dc = ws + (2 * d + 1) * sf;
    float* wtd = wt + (int64_t)d * 4 * H * H;
    // B operand of dh = dG . W_hh: rows = hidden units (N = H), K = gate rows (4H)  ->  planes of W_hh^T [H, 4H]
    int rc = transpose_launch(w_hh[d], wtd, 4 * H, H, sd);
    if (rc) return rc;
    if ((rc = tc16::split_planes(wtd, 4 * H, H, 4 * H, 1.f, wpl, sd, true))) return rc;
    planes_bwd_init_kernel<<<ceil_div(sf > RP ? sf : RP, 256), 256, 0, sd>>>(d_hn ? d_hn + d * dirn : nullptr, d_cn ? d_cn + d * dirn : nullptr,
                                                                             ldn, carry, dc, dh_rec, inv_s, amax + d, B, RP, H);
    DVAE_LAUNCH_CHECK();
    const int nsteps = T + ((d_h0 || d_c0) ? 1 : 0);
    for (int s = 0; s < nsteps; ++s) {
      const bool final_ = s == T;
      const int t = d == 0 ? T - 1 - s : s;
      if (s > 0) {      // dh_rec [B,H] (+)= (dG_{t_next} rows * 2^e) . W_hh, rows unscaled by inv_s in the epilogue
        rc = tc16::linear_planes(apl, wpl, dh_rec, H, B, H, 4 * H, nullptr, 0.f, 0, 1.f, 1.f, inv_s, true, ms, sd);
        if (rc) return rc;
      }
      if (final_) {
        planes_bwd_final_kernel<<<ceil_div(sf, 256), 256, 0, sd>>>(carry, dc, dh_rec, d_h0 ? d_h0 + d * dird0 : nullptr,
                                                                   d_c0 ? d_c0 + d * dird0 : nullptr, ldd0, B, H);
        DVAE_LAUNCH_CHECK();
        break;
      }
      const int t_prev = d == 0 ? t - 1 : t + 1;      // step that ran before t in the forward traversal
      CellBwdArgs a;
      a.gates_t = gates + d * slab + (int64_t)t * B * 4 * H;
      a.cs_t = cs + ((int64_t)d * T + t) * sf;
      if (t_prev >= 0 && t_prev < T) { a.c_prev = cs + ((int64_t)d * T + t_prev) * sf; a.ldcp = H; }
      else { a.c_prev = c0 ? c0 + d * dir0 : nullptr; a.ldcp = ld0; }
      a.d_hs_t = d_hs ? d_hs + (int64_t)t * B * lddhs + d * H : nullptr; a.lddhs = lddhs;
      a.carry = carry; a.dc = dc; a.dh_rec = dh_rec; a.lengths = lengths; a.planes = apl; a.inv_s = inv_s; a.amax = amax + d;
      a.t = t; a.B = B; a.H = H; a.use_rec = s > 0 ? 1 : 0;
      lstm_cell_bwd_kernel<<<RP, kCellThreads, 0, sd>>>(a);
      DVAE_LAUNCH_CHECK();
    }
  }
  return fork.join();
}

}  // namespace dvae
