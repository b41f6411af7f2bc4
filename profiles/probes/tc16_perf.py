"""fp16-split tcgen05 GEMM (dvae_tc16_linear) vs the 3xTF32 kernel (dvae_tc_linear) on the cfg-2 train-step shapes,
plus the fused vocab-CE forward / backward entry points (which pick the fp16-split kernel by default;
DVAE_GEMM_IMPL=tf32 selects the 3xTF32 one).  Warm clocks, CUDA events, TFLOP/s = fp32-equivalent 2*M*N*K."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
st = L.stream_ptr()


def bench(fn, iters=50):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


x = torch.randn(4096, 4096, device="cuda")
for _ in range(200): x @ x          # warm clocks
shapes = [("lstm input proj fwd", 2816, 1024, 256, 0, 0), ("lstm input proj l1", 2816, 1024, 512, 0, 0), ("dx = dG W_ih", 2688, 256, 1024, 0, 1),
          ("dW_ih = dG^T x", 1024, 256, 2688, 1, 1), ("dh = P W_out", 2688, 256, 4608, 0, 1), ("dW_out = P^T h", 4608, 256, 2688, 1, 1),
          ("logits-like", 2688, 10000, 256, 0, 0), ("square 4096", 4096, 4096, 4096, 0, 0)]
for (name, M, N, K, ta, tb) in shapes:
    A = torch.randn((K, M) if ta else (M, K), device="cuda"); B = torch.randn((K, N) if tb else (N, K), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    fl = 2.0 * M * N * K
    u16 = bench(lambda: lib.dvae_tc16_linear(L.ptr(A), A.stride(0), ta, L.ptr(B), B.stride(0), tb, L.ptr(C), N, M, N, K, None, None, 0.0, 0, 1.0, 1.0, None, None, st))
    u3 = bench(lambda: lib.dvae_tc_linear(L.ptr(A), A.stride(0), ta, L.ptr(B), B.stride(0), tb, L.ptr(C), N, M, N, K, None, None, 0.0, 0, 3, st))
    print(f"{name:22s} M={M:5d} N={N:5d} K={K:5d} ta={ta} tb={tb}: fp16-split {u16:7.1f} us {fl / u16 / 1e6:6.1f} TF/s | 3xTF32 {u3:7.1f} us {fl / u3 / 1e6:6.1f} TF/s", flush=True)

# fused vocab-CE at cfg 2: N = 21*128 positions, H = 256, V = 10000
T1, Bt, H, V = 21, 128, 256, 10000
N = T1 * Bt
h = torch.randn(T1, Bt, H, device="cuda") * 0.5
w = torch.randn(V, H, device="cuda") * 0.05
bias = torch.zeros(V, device="cuda")
tg = torch.randint(4, V, (Bt, T1 + 1), device="cuda")
ln = torch.full((Bt,), T1 + 1, device="cuda", dtype=torch.int64)
lse, nll = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
am = torch.zeros(N, device="cuda", dtype=torch.int32)
loss = torch.zeros(1, device="cuda")
ws = torch.zeros(lib.dvae_vocab_ce_ws_floats(N, V, H), device="cuda")
wsb = torch.zeros(lib.dvae_vocab_ce_bwd_ws_floats(N, V, H), device="cuda")
dh, dw, db = torch.zeros(N, H, device="cuda"), torch.zeros(V, H, device="cuda"), torch.zeros(V, device="cuda")
fwd = lambda: lib.dvae_vocab_ce_fwd(L.ptr(h), H, T1, Bt, H, V, L.ptr(w), L.ptr(bias), L.ptr(tg), tg.stride(0), L.ptr(ln), 2, L.ptr(lse), L.ptr(nll), L.ptr(am), L.ptr(loss), L.ptr(ws), st)
bwd = lambda: lib.dvae_vocab_ce_bwd(L.ptr(h), H, T1, Bt, H, V, L.ptr(w), L.ptr(bias), L.ptr(tg), tg.stride(0), L.ptr(ln), L.ptr(lse), None, L.ptr(dh), H, L.ptr(dw), L.ptr(db), None, L.ptr(wsb), st)
uf, ub = bench(fwd), bench(bwd)
fl = 2.0 * N * H * V
print(f"vocab-CE forward  (N={N}, H={H}, V={V}): {uf:7.1f} us  {fl / uf / 1e6:6.1f} TF/s algorithmic")
print(f"vocab-CE backward (recompute + dh + dW + db): {ub:7.1f} us  {3 * fl / ub / 1e6:6.1f} TF/s algorithmic (6*N*H*V)")
