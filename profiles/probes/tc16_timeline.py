"""In-kernel milestone timestamps (globaltimer, ns) of CTA (0,0,0) of the fp16-split tcgen05 GEMM (tc_gemm16.cu)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
os.environ["DVAE_TC_DBG"] = hex(dbg.data_ptr())
x = torch.randn(4096, 4096, device="cuda")
for _ in range(200): x @ x
names = {0: "entry", 2: "producer: k-block 12 begin", 3: "producer: ring slot free", 4: "producer: split + stores issued", 5: "producer: fence + arrive",
         6: "producer: k-block 13 done", 8: "mma: waits for k-block 12", 9: "mma: k-block 12 operands ready", 10: "mma: k-block 12 issued + committed",
         11: "mma: k-block 13 operands ready", 14: "epilogue: tile 0 accumulators ready", 15: "epilogue: tile 0 done", 1: "exit"}
for (M, N, K, ta, tb) in ((4096, 4096, 2048, 0, 0), (1024, 1024, 2048, 0, 0), (512, 512, 2048, 0, 0), (5120, 256, 2688, 1, 1), (2688, 256, 5120, 0, 1)):
    A = torch.randn((K, M) if ta else (M, K), device="cuda"); B = torch.randn((K, N) if tb else (N, K), device="cuda"); C = torch.zeros(M, N, device="cuda")
    for it in range(5):
        dbg.zero_()
        st_ev, en_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st_ev.record()
        lib.dvae_tc16_linear(L.ptr(A), A.stride(0), ta, L.ptr(B), B.stride(0), tb, L.ptr(C), N, M, N, K, None, None, 0.0, int(os.environ.get('PROBE_ACT', '0')), 1.0, 1.0, None, None, L.stream_ptr())
        en_ev.record(); torch.cuda.synchronize()
    t = dbg.cpu().tolist()
    print(f"M={M} N={N} K={K} ta={ta} tb={tb}: event time {st_ev.elapsed_time(en_ev) * 1e3:.1f} us")
    for i in sorted(names, key=lambda i: t[i]):
        if t[i]:
            print(f"   {names[i]:40s} +{(t[i] - t[0]) / 1e3:8.2f} us")
