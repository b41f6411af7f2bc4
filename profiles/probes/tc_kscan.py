"""Device time of one full wave (148 tiles) of the tcgen05 GEMM vs K: separates fixed per-tile cost from per-k-block cost."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
st = L.stream_ptr()
M, N = 128 * 37, 512
for passes in (3, 1):
    for K in (32, 64, 128, 256, 512, 1024, 2048):
        A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.zeros(M, N, device="cuda")
        for _ in range(2):
            lib.dvae_tc_linear(L.ptr(A), K, 0, L.ptr(B), K, 0, L.ptr(C), N, M, N, K, None, None, 0.0, 0, passes, st)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                lib.dvae_tc_linear(L.ptr(A), K, 0, L.ptr(B), K, 0, L.ptr(C), N, M, N, K, None, None, 0.0, 0, passes, st)
            torch.cuda.synchronize()
        ts = [e.device_time for e in prof.events() if "tc_gemm" in e.name]
        print(f"passes={passes} K={K:5d} kblocks={K // 32:3d}: {sum(ts) / len(ts):8.2f} us per wave", flush=True)
