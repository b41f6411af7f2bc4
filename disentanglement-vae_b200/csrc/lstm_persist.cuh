// Argument blocks and entry points of the persistent (cluster, SMEM-resident W_hh) LSTM kernels.
#pragma once
#include "common.cuh"

namespace dvae {

struct PersistFwdArgs {
  const float* w_hh[2];
  float* gates; float* cs; float* hs; int64_t ldhs;
  const float* h0; const float* c0; int64_t ld0, dir0;
  float* hstate;                 // [2][D][B][H] carried-h exchange buffer
  float* hn; float* cn; int64_t ldn, dirn;
  const int64_t* lengths;
  int T, B, D, n_slices, d_off;
  unsigned long long* dbg;   // optional per-step milestone timestamps (probes)
};

struct PersistBwdArgs {
  const float* w_hh[2];
  float* gates;                  // post-activation gates in, dG out
  const float* cs;
  const float* c0; int64_t ld0, dir0;
  const float* d_hs; int64_t lddhs;
  const float* d_hn; const float* d_cn; int64_t ldn, dirn;
  float* d_h0; float* d_c0; int64_t ldd0, dird0;
  const int64_t* lengths;
  int T, B, D, n_slices, d_off;
  uint32_t* amax_out;            // optional: [grid size] per-CTA bit patterns of max |dG| (operand scale of the dense gradient GEMMs)
  float* zero_buf; int64_t zero_n4;   // optional: float4s cleared by the kernel on its way in (d_x, which a split-K GEMM accumulates into)
};

// true when the persistent kernels can take this call (H in {64,128,256}, 16-byte aligned buffers)
bool persist_supported(int B, int H, int D, const void* const* ptrs, int nptr, const int64_t* lds, int nld);
int persist_fwd(int H, const PersistFwdArgs& a, cudaStream_t st);
int persist_bwd(int H, const PersistBwdArgs& a, cudaStream_t st);

// tcgen05 variant (lstm_tc.cu): H = 256, fp16 hi/lo split operands resident in SMEM, accumulators in TMEM.
// DVAE_LSTM_IMPL=simt keeps the fp32 SIMT persistent kernels (A/B tests).
bool tc_lstm_supported(int H);
int tc_lstm_fwd(const PersistFwdArgs& a, cudaStream_t st);
int tc_lstm_bwd(const PersistBwdArgs& a, cudaStream_t st);
int tc_lstm_bwd_ctas(int B, int D);          // grid size of that launch (= per-CTA max |dG| entries written)
// entries reserved per max |dG| slot set: the largest grid the backward kernel can have for (B, D) (one row group per cluster)
inline int amax_slot_entries(int B, int D) { const int n = 8 * ((B + 15) / 16) * D; return n < 128 ? 128 : n; }


// Large-H path (lstm_planes.cu): per-step tensor-core GEMMs over fp16 operand planes + element-wise cell kernels, for hidden
// sizes the cluster-resident kernels do not take (H >= 128, H % 32 == 0; e.g. 512, 1024).
bool planes_lstm_supported(int B, int H, int D, const void* const* ptrs, int nptr, const int64_t* lds, int nld);
int64_t planes_lstm_ws_floats(int B, int H, int D);
int planes_lstm_fwd(int T, int B, int H, int D, const float* const* w_hh, const float* h0, const float* c0, int64_t ld0,
                    int64_t dir0, const int64_t* lengths, float* hs, int64_t ldhs, float* hn, float* cn, int64_t ldn,
                    int64_t dirn, float* gates, float* cs, float* ws, float* pws, cudaStream_t st);
int planes_lstm_bwd(int T, int B, int H, int D, const float* const* w_hh, const float* c0, int64_t ld0, int64_t dir0,
                    const int64_t* lengths, float* gates, const float* cs, const float* d_hs, int64_t lddhs, const float* d_hn,
                    const float* d_cn, int64_t ldn, int64_t dirn, float* d_h0, float* d_c0, int64_t ldd0, int64_t dird0,
                    float* ws, float* wt, float* pws, uint32_t* amax_slots, cudaStream_t st);

// Grid-resident forward recurrence (lstm_resident.cu): H = 512 / 1024, B <= 128 -- W_hh tiled over the grid's tensor memory,
// one launch per direction for all T steps.
bool resident_lstm_supported(int B, int H);
int64_t resident_lstm_ws_floats(int H);
int resident_lstm_fwd(int T, int B, int H, int D, const float* const* w_hh, const float* h0, const float* c0, int64_t ld0,
                      int64_t dir0, const int64_t* lengths, float* hs, int64_t ldhs, float* hn, float* cn, int64_t ldn,
                      int64_t dirn, float* gates, float* cs, float* pws, cudaStream_t st);

}  // namespace dvae
