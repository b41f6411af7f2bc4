"""Tile-level timeline of the softmax-gradient kernel (tc16 mode 2, cfg-2 shape) inside dvae_vocab_ce_bwd: when the
epilogue warps of CTA (0,0,0) see each tile's accumulators and when they are done with them (globaltimer ns)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
os.environ["DVAE_TC_DBG"] = hex(dbg.data_ptr())
os.environ["DVAE_TC_DBG_MODE"] = "2"
x = torch.randn(4096, 4096, device="cuda")
for _ in range(200): x @ x
T1, Bt, H, V = 21, 128, 256, 10000
N = T1 * Bt
h = torch.randn(T1, Bt, H, device="cuda") * 0.5; w = torch.randn(V, H, device="cuda") * 0.05; bias = torch.zeros(V, device="cuda")
tg = torch.randint(4, V, (Bt, T1 + 1), device="cuda"); ln = torch.full((Bt,), T1 + 1, device="cuda", dtype=torch.int64)
lse, nll = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
am = torch.zeros(N, device="cuda", dtype=torch.int32); loss = torch.zeros(1, device="cuda")
ws = torch.zeros(lib.dvae_vocab_ce_ws_floats(N, V, H), device="cuda")
wsb = torch.zeros(lib.dvae_vocab_ce_bwd_ws_floats(N, V, H), device="cuda")
dh, dw, db = torch.zeros(N, H, device="cuda"), torch.zeros(V, H, device="cuda"), torch.zeros(V, device="cuda")
st = L.stream_ptr()
L.check(lib.dvae_vocab_ce_fwd(L.ptr(h), H, T1, Bt, H, V, L.ptr(w), L.ptr(bias), L.ptr(tg), tg.stride(0), L.ptr(ln), 2, L.ptr(lse), L.ptr(nll),
                              L.ptr(am), L.ptr(loss), L.ptr(ws), st), "ce")
for _ in range(5):
    dbg.zero_()
    L.check(lib.dvae_vocab_ce_bwd(L.ptr(h), H, T1, Bt, H, V, L.ptr(w), L.ptr(bias), L.ptr(tg), tg.stride(0), L.ptr(ln), L.ptr(lse), None,
                                  L.ptr(dh), H, L.ptr(dw), L.ptr(db), None, L.ptr(wsb), st), "ce bwd")
    torch.cuda.synchronize()
t = dbg.cpu().tolist()
print("softmax-gradient kernel (last vocabulary chunk): entry +0.00 us; exit +%.2f us" % ((t[1] - t[0]) / 1e3))
for k in range(4):
    print(f"  tile {k}: accumulators ready +{(t[14 + 2 * k] - t[0]) / 1e3:6.2f} us, epilogue done +{(t[15 + 2 * k] - t[0]) / 1e3:6.2f} us")
if t[104]:      # probe build: inside warp 0's epilogue of tile 2 (two 32-column chunks)
    for c in range(2):
        a, b, d = (t[104 + 3 * c + i] for i in range(3))
        print(f"  tile 2 chunk {c}: start +{(a - t[104]) / 1e3:5.2f} us | TMEM loads + exp + transpose stores {(b - a) / 1e3:5.2f} us | global stores {(d - b) / 1e3:5.2f} us")
