"""In-kernel milestone timestamps (globaltimer, ns) of CTA (0,0,0) of the tcgen05 GEMM."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
dbg = torch.zeros(32, dtype=torch.int64, device="cuda")
os.environ["DVAE_TC_DBG"] = hex(dbg.data_ptr())
names = ["entry", "setup done", "first TMA issued", "first raw landed", "first split done (MMA sees)", "last MMA committed", "epilogue sees acc", "epilogue done", "exit"] + [f"chunk{c} {w}" for c in range(4) for w in ("tmem loaded", "staged in smem", "stored")]
for (M, N, K) in ((128 * 37, 512, 256),):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.zeros(M, N, device="cuda")
    for it in range(3):
        dbg.zero_()
        st_ev, en_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st_ev.record()
        lib.dvae_tc_linear(L.ptr(A), K, 0, L.ptr(B), K, 0, L.ptr(C), N, M, N, K, None, None, 0.0, 0, 3, L.stream_ptr())
        en_ev.record(); torch.cuda.synchronize()
    t = dbg.cpu().tolist()
    print(f"M={M} N={N} K={K}: event time {st_ev.elapsed_time(en_ev) * 1e3:.1f} us")
    for i, n in enumerate(names):
        print(f"   {n:32s} +{(t[i] - t[0]) / 1e3:8.2f} us")
