"""Tile-level timeline of the fused vocab-CE forward kernel (tc16 mode 1, cfg-2 shape): when the epilogue warps of
CTA (0,0,0) see each tile's accumulators and when they are done with them (globaltimer ns)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
os.environ["DVAE_TC_DBG"] = hex(dbg.data_ptr())
x = torch.randn(4096, 4096, device="cuda")
for _ in range(200): x @ x
T1, Bt, H, V = 21, 128, 256, 10000
N = T1 * Bt
h = torch.randn(T1, Bt, H, device="cuda") * 0.5; w = torch.randn(V, H, device="cuda") * 0.05; bias = torch.zeros(V, device="cuda")
tg = torch.randint(4, V, (Bt, T1 + 1), device="cuda"); ln = torch.full((Bt,), T1 + 1, device="cuda", dtype=torch.int64)
lse, nll = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
am = torch.zeros(N, device="cuda", dtype=torch.int32); loss = torch.zeros(1, device="cuda")
ws = torch.zeros(lib.dvae_vocab_ce_ws_floats(N, V, H), device="cuda")
for _ in range(5):
    dbg.zero_()
    L.check(lib.dvae_vocab_ce_fwd(L.ptr(h), H, T1, Bt, H, V, L.ptr(w), L.ptr(bias), L.ptr(tg), tg.stride(0), L.ptr(ln), 2, L.ptr(lse), L.ptr(nll),
                                  L.ptr(am), L.ptr(loss), L.ptr(ws), L.stream_ptr()), "ce")
    torch.cuda.synchronize()
t = dbg.cpu().tolist()
print("entry +0.00 us; exit +%.2f us" % ((t[1] - t[0]) / 1e3))
for k in range(4):
    print(f"  tile {k}: accumulators ready +{(t[14 + 2 * k] - t[0]) / 1e3:6.2f} us, epilogue done +{(t[15 + 2 * k] - t[0]) / 1e3:6.2f} us")
if t[96]:       # probe build (-DDVAE_TC16_PROBES): inside warp 0's epilogue of tile 2, and the per-k-block marks of tile 1
    names = ["accumulators seen", "bias staged", "hh0: before tmem loads", "hh0: loads done", "hh0: math done", "hh1: before tmem loads",
             "hh1: loads done", "hh1: math done"]
    for i, n in enumerate(names):
        print(f"  tile 2 epilogue (warp 0): {n:24s} +{(t[96 + i] - t[96]) / 1e3:5.2f} us")
    for k in range(8):
        print(f"  tile 1 k-block {k}: producer waits for the slot from +{(t[32 + 2 * k] - t[0]) / 1e3:6.2f}, issues TMA +{(t[33 + 2 * k] - t[0]) / 1e3:6.2f} | "
              f"mma warp waits from +{(t[48 + 2 * k] - t[0]) / 1e3:6.2f}, operands ready +{(t[49 + 2 * k] - t[0]) / 1e3:6.2f} us")
