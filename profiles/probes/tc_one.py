"""Run one tcgen05 GEMM shape a few times (target for ncu)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
M, N, K = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 4096, 256)))
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.zeros(M, N, device="cuda")
for _ in range(3):
    L.check(lib.dvae_tc_linear(L.ptr(A), K, 0, L.ptr(B), K, 0, L.ptr(C), N, M, N, K, None, None, 0.0, 0, 3, L.stream_ptr()), "tc")
torch.cuda.synchronize()
print("ok", float(C.abs().mean()))
