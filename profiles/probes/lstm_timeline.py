"""Persistent LSTM kernels at the cfg-2 shape (B=128, H=256): per-step cost of forward and backward (slope between
T=2 and T=22 whole-layer calls, warm clocks, CUDA events) and, for the forward kernel, the milestones inside one step
(globaltimer ns, CTA 0 / thread 0).  DVAE_LSTM_IMPL=simt selects the fp32 SIMT persistent kernels."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
dvae = importlib.import_module("disentanglement-vae_b200"); L = dvae._lib; lib = L.load()
B, I, H = 128, 256, 256
D = int(os.environ.get("PROBE_D", "1"))
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
pa = lambda ts: L.ptr_array(ts)


def make(T):
    g = torch.Generator(device="cuda").manual_seed(1)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    W = dict(w_ih=[r(4 * H, I) * 0.05 for _ in range(D)], w_hh=[r(4 * H, H) * 0.05 for _ in range(D)],
             b_ih=[torch.zeros(4 * H, device="cuda") for _ in range(D)], b_hh=[torch.zeros(4 * H, device="cuda") for _ in range(D)])
    G = {k: [torch.zeros_like(t) for t in v] for k, v in W.items()}
    bufs = dict(x=r(T, B, I), hs=torch.zeros(T, B, D * H, device="cuda"), gates=torch.zeros(D, T, B, 4 * H, device="cuda"),
                cs=torch.zeros(D, T, B, H, device="cuda"), ws=torch.zeros(lib.dvae_lstm_state_ws_floats(B, H, D), device="cuda"),
                d_hs=r(T, B, D * H) * 1e-3, d_x=torch.zeros(T, B, I, device="cuda"),
                lengths=torch.randint(min(3, T), T + 1, (B,), device="cuda", dtype=torch.int64, generator=g))
    return W, G, bufs


def fwd(T, W, G, b):
    L.check(lib.dvae_lstm_seq_fwd(L.ptr(b["x"]), I, T, B, I, H, D, pa(W["w_ih"]), pa(W["w_hh"]), pa(W["b_ih"]), pa(W["b_hh"]), None, None, 0, 0,
                                  L.ptr(b["lengths"]), L.ptr(b["hs"]), D * H, None, None, 0, 0, L.ptr(b["gates"]), L.ptr(b["cs"]), L.ptr(b["ws"]),
                                  L.stream_ptr()), "fwd")


def bwd(T, W, G, b):
    L.check(lib.dvae_lstm_seq_bwd(L.ptr(b["x"]), I, T, B, I, H, D, pa(W["w_ih"]), pa(W["w_hh"]), None, None, 0, 0, L.ptr(b["lengths"]), L.ptr(b["hs"]),
                                  D * H, L.ptr(b["gates"]), L.ptr(b["cs"]), L.ptr(b["d_hs"]), D * H, None, None, 0, 0, L.ptr(b["d_x"]), I,
                                  pa(G["w_ih"]), pa(G["w_hh"]), pa(G["b_ih"]), pa(G["b_hh"]), None, None, 0, 0, L.ptr(b["ws"]), L.stream_ptr()), "bwd")


def timed(fn, n=100):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


res = {}
for T in (2, 22):
    W, G, b = make(T)
    for _ in range(300):            # warm clocks
        fwd(T, W, G, b)
    res[("fwd", T)] = timed(lambda: fwd(T, W, G, b))
    res[("fwd+bwd", T)] = timed(lambda: (fwd(T, W, G, b), bwd(T, W, G, b)))
for k in ("fwd", "fwd+bwd"):
    print(f"{k:8s} layer call D={D}: T=2 {res[(k, 2)]:7.1f} us, T=22 {res[(k, 22)]:7.1f} us -> {(res[(k, 22)] - res[(k, 2)]) / 20:5.2f} us per step (incl. hoisted GEMM share)")
os.environ["DVAE_LSTM_DBG"] = hex(dbg.data_ptr())
W, G, b = make(22)
for _ in range(50):
    fwd(22, W, G, b)
torch.cuda.synchronize()
t = dbg.cpu().tolist()
if os.environ.get("DVAE_LSTM_IMPL") == "simt":
    for i, n in enumerate(["step start", "h staged (L2 -> smem)", "mini-GEMM done", "gates/epilogue done", "cluster barrier passed"]):
        print(f"  {n:28s} +{(t[i] - t[0]) / 1e3:6.2f} us")
else:   # tcgen05 kernel (lstm_tc.cu): marks 0/1/9 are per launch, 2..8 belong to step 6
    print(f"  setup (W split, TMEM alloc, cluster sync) {(t[1] - t[0]) / 1e3:6.2f} us; whole kernel {(t[9] - t[0]) / 1e3:6.2f} us")
    for i, n in zip(range(2, 9), ["step start", "state slices landed (mbarrier)", "48 MMAs issued + commit", "MMAs complete (mbarrier)",
                                  "gates activated (tcgen05.ld, sts, bar)", "cell update + stores", "state published"]):
        print(f"  {n:46s} +{(t[i] - t[2]) / 1e3:6.2f} us")
