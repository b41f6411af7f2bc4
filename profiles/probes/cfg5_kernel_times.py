"""Per-kernel device time (by launch grid) of one cfg-5 resample (inference.ConsistencyEvaluator, eager launches) via CUPTI.
Analysis aid only -- never a bench value."""
import collections, importlib, json, os, sys
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
ev = inf.ConsistencyEvaluator(vae, B.BATCH, B.SEQ_T, use_graph=os.environ.get("TRACE_GRAPH", "0") == "1")
ev.encode_once(torch.from_numpy(X), torch.from_numpy(L))
ev.resample(2)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ev.resample(2)
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "cfg5_trace.json")
prof.export_chrome_trace(path)
agg = collections.defaultdict(lambda: [0, 0.0])
evs = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
os.remove(path)
for e in evs:
    nm = e["name"].split("(")[0].replace("void ", "").replace("dvae::", "").replace("(anonymous namespace)::", "")[:60]
    a = agg[f"{nm} {e['args'].get('grid', '')}"]
    a[0] += 1; a[1] += e["dur"]
tot = sum(v[1] for v in agg.values())
span = (max(e["ts"] + e["dur"] for e in evs) - min(e["ts"] for e in evs)) / 2
print(f"# cfg5: per resample {tot / 2:.1f} us summed kernel time, {span:.1f} us span (eager)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t / 2:10.1f} us {100 * t / tot:5.1f}% x{c // 2:4d}  {k}")

# the largest idle gaps between consecutive kernels of the second resample (graph replays expose launch / dependency latencies)
evs.sort(key=lambda e: e["ts"])
half = evs[len(evs) // 2:]
end, gaps = half[0]["ts"], []
for e in half:
    g = e["ts"] - end
    if g > 0:
        gaps.append((g, e["name"][:50]))
    end = max(end, e["ts"] + e["dur"])
print(f"# second resample: {len(half)} kernels, idle {sum(g for g, _ in gaps):.1f} us in {len(gaps)} gaps; largest:")
for g, n in sorted(gaps, reverse=True)[:12]:
    print(f"   {g:8.1f} us before {n}")
