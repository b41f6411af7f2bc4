"""Launch sequence for `ncu --set full` captures of the three hot kernels at the cfg-2 shape:
   3 x vocab-CE forward (tc_gemm_kernel mode 1 + finalize), then one decoder-like LSTM layer forward + backward
   (input-projection GEMM, lstm_tc_fwd_kernel<1>, lstm_tc_bwd_kernel, dense-gradient GEMMs), then one bidirectional
   encoder-like layer forward (lstm_tc_fwd_kernel<2>).
   ncu -k regex:'tc_gemm_kernel|lstm_tc' -c 12 --set full --import-source on ... python profiles/probes/ncu_targets.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
st = L.stream_ptr()
os.environ["DVAE_FORK"] = "0"
T1, Bt, H, V, I = 21, 128, 256, 10000, 256
N = T1 * Bt
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, device="cuda", generator=g)
h = r(T1, Bt, H) * 0.5; w = r(V, H) * 0.05; bias = torch.zeros(V, device="cuda")
tg = torch.randint(4, V, (Bt, T1 + 1), device="cuda"); ln = torch.full((Bt,), T1 + 1, device="cuda", dtype=torch.int64)
lse, nll = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
am = torch.zeros(N, device="cuda", dtype=torch.int32); loss = torch.zeros(1, device="cuda")
ws = torch.zeros(lib.dvae_vocab_ce_ws_floats(N, V, H), device="cuda")
for _ in range(3):
    L.check(lib.dvae_vocab_ce_fwd(L.ptr(h), H, T1, Bt, H, V, L.ptr(w), L.ptr(bias), L.ptr(tg), tg.stride(0), L.ptr(ln), 2, L.ptr(lse), L.ptr(nll),
                                  L.ptr(am), L.ptr(loss), L.ptr(ws), st), "ce")
torch.cuda.synchronize()
pa = lambda ts: L.ptr_array(ts)
for D, T in ((1, T1), (2, 22)):
    W = dict(w_ih=[r(4 * H, I) * 0.05 for _ in range(D)], w_hh=[r(4 * H, H) * 0.05 for _ in range(D)],
             b_ih=[torch.zeros(4 * H, device="cuda") for _ in range(D)], b_hh=[torch.zeros(4 * H, device="cuda") for _ in range(D)])
    G = {k: [torch.zeros_like(t) for t in v] for k, v in W.items()}
    x = r(T, Bt, I); hs = torch.zeros(T, Bt, D * H, device="cuda"); gates = torch.zeros(D, T, Bt, 4 * H, device="cuda")
    cs = torch.zeros(D, T, Bt, H, device="cuda"); lws = torch.zeros(lib.dvae_lstm_state_ws_floats(Bt, H, D), device="cuda")
    d_hs = r(T, Bt, D * H) * 1e-3; d_x = torch.zeros(T, Bt, I, device="cuda")
    lens = torch.randint(3, T + 1, (Bt,), device="cuda", dtype=torch.int64, generator=g)
    L.check(lib.dvae_lstm_seq_fwd(L.ptr(x), I, T, Bt, I, H, D, pa(W["w_ih"]), pa(W["w_hh"]), pa(W["b_ih"]), pa(W["b_hh"]), None, None, 0, 0,
                                  L.ptr(lens), L.ptr(hs), D * H, None, None, 0, 0, L.ptr(gates), L.ptr(cs), L.ptr(lws), st), "fwd")
    if D == 1:
        L.check(lib.dvae_lstm_seq_bwd(L.ptr(x), I, T, Bt, I, H, D, pa(W["w_ih"]), pa(W["w_hh"]), None, None, 0, 0, L.ptr(lens), L.ptr(hs), D * H,
                                      L.ptr(gates), L.ptr(cs), L.ptr(d_hs), D * H, None, None, 0, 0, L.ptr(d_x), I, pa(G["w_ih"]), pa(G["w_hh"]),
                                      pa(G["b_ih"]), pa(G["b_hh"]), None, None, 0, 0, L.ptr(lws), st), "bwd")
    torch.cuda.synchronize()
print("ok")
