"""Per-step milestones (globaltimer ns) of CTA 0 of the persistent LSTM forward kernel at the cfg-2 shape."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["DVAE_LSTM_DBG"] = hex(dbg.data_ptr())
T, B, I, H, D = 22, 128, 256, 256, 1
x = torch.randn(T, B, I, device="cuda")
W = [torch.randn(4 * H, I, device="cuda") * 0.05, torch.randn(4 * H, H, device="cuda") * 0.05, torch.zeros(4 * H, device="cuda"), torch.zeros(4 * H, device="cuda")]
hs = torch.zeros(T, B, D * H, device="cuda"); gates = torch.zeros(D, T, B, 4 * H, device="cuda"); cs = torch.zeros(D, T, B, H, device="cuda")
ws = torch.zeros(lib.dvae_lstm_state_ws_floats(B, H, D), device="cuda")
pa = lambda t: L.ptr_array([t])
for it in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.check(lib.dvae_lstm_seq_fwd(L.ptr(x), I, T, B, I, H, D, pa(W[0]), pa(W[1]), pa(W[2]), pa(W[3]), None, None, 0, 0, None, L.ptr(hs), D * H,
                                  None, None, 0, 0, L.ptr(gates), L.ptr(cs), L.ptr(ws), L.stream_ptr()), "lstm")
    b.record(); torch.cuda.synchronize()
t = dbg.cpu().tolist()
print(f"whole layer call (input GEMM + {T} steps): {a.elapsed_time(b) * 1e3:.1f} us")
for i, n in enumerate(["step start", "h staged (L2 -> smem)", "mini-GEMM done", "gates/epilogue done", "cluster barrier passed"]):
    print(f"  {n:28s} +{(t[i] - t[0]) / 1e3:6.2f} us")
