"""Where a cfg-5 resample's wall time goes: graph replays back to back vs with the per-resample result copies vs eager."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import __graft_entry__ as ge
dvae = ge.build()
inf = importlib.import_module("disentanglement-vae_b200.inference")
B.select_workload("cfg5")
dev = torch.device("cuda", 0)
dvae.set_seed(10)
vae = dvae.build_vae(B.CFG2, B.VOCAB, None, B.LABELS, dev, B.SOS, B.EOS); vae.train()
rng = np.random.default_rng(5)
X, L, _ = B.synth_batch(rng, B.BATCH)
ev = inf.ConsistencyEvaluator(vae, B.BATCH, B.SEQ_T, use_graph=True)
ev.encode_once(torch.from_numpy(X), torch.from_numpy(L))
ev.resample(3)
torch.cuda.synchronize()
g = ev._graph[False]
def timed(fn, n=10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (t1 - t0) * 1e3 / n, (time.perf_counter() - t0) * 1e3 / n
print("graph replay only      : device %.3f ms, host enqueue %.3f ms, wall %.3f ms per resample" % timed(g.replay))
out = ev.resample(1)
print("resample(10) API       : device %.3f ms, host enqueue %.3f ms, wall %.3f ms per resample" % timed(lambda: ev.resample(10), 2) )
ev.use_graph = False
ev._planes_on(True)
print("eager body             : device %.3f ms, host enqueue %.3f ms, wall %.3f ms per resample" % timed(lambda: ev._body(False), 5))
ev._planes_off()
