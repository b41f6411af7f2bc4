"""Microbenchmark of the tcgen05 GEMM (dvae_tc_linear) vs the SIMT GEMM: TFLOP/s by shape and pass count."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
st = L.stream_ptr()
def bench(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
shapes = [(4096, 4096, 256, 0, 0), (4096, 4096, 2048, 0, 0), (2688, 1024, 256, 0, 0), (2816, 1024, 512, 0, 0), (2688, 10000, 256, 0, 0),
          (2688, 256, 1024, 0, 1), (1024, 256, 2688, 1, 1), (10000, 256, 2688, 1, 1), (2688, 256, 10000, 0, 1), (128, 128, 4096, 0, 0), (128, 128, 256, 0, 0)]
for (M, N, K, ta, tb) in shapes:
    A = torch.randn((K, M) if ta else (M, K), device="cuda"); B = torch.randn((K, N) if tb else (N, K), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    fl = 2.0 * M * N * K
    row = f"M={M:6d} N={N:6d} K={K:6d} ta={ta} tb={tb}:"
    for passes in (3, 1):
        us = bench(lambda: lib.dvae_tc_linear(L.ptr(A), A.stride(0), ta, L.ptr(B), B.stride(0), tb, L.ptr(C), N, M, N, K, None, None, 0.0, 0, passes, st))
        row += f"  tc{passes}: {us:8.1f} us {fl / us / 1e6:7.1f} TF/s"
    os.environ["DVAE_GEMM_IMPL"] = "simt"
    us = bench(lambda: lib.dvae_linear(L.ptr(A), A.stride(0), ta, L.ptr(B), B.stride(0), tb, L.ptr(C), N, M, N, K, None, None, 0.0, 0, st), 5)
    del os.environ["DVAE_GEMM_IMPL"]
    row += f"  simt: {us:8.1f} us {fl / us / 1e6:7.1f} TF/s"
    print(row, flush=True)
