"""One vocab-CE forward and two backward calls at the cfg-2 shape (ncu target: tc16 launches 0 = CE forward,
1/4/7/10 = softmax-gradient chunks, the others dh / dW GEMMs)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
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
for _ in range(2):
    L.check(lib.dvae_vocab_ce_bwd(L.ptr(h), H, T1, Bt, H, V, L.ptr(w), L.ptr(bias), L.ptr(tg), tg.stride(0), L.ptr(ln), L.ptr(lse), None,
                                  L.ptr(dh), H, L.ptr(dw), L.ptr(db), None, L.ptr(wsb), st), "ce bwd")
torch.cuda.synchronize()
