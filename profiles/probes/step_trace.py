"""Kernel timeline of one cfg-2 train step (eager launches, fork/join side streams) from CUPTI via torch.profiler:
start / duration / stream of every kernel, and the gaps on the critical path.  Not a bench: tracing adds overhead."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import __graft_entry__ as ge
dvae = ge.build()
engine_mod = importlib.import_module("disentanglement-vae_b200.engine")
dev = torch.device("cuda", 0)
B.select_workload("cfg2")
dvae.set_seed(10)
vae = dvae.build_vae(B.CFG2, B.VOCAB, None, B.LABELS, dev, B.SOS, B.EOS); vae.train()
use_graph = os.environ.get("TRACE_GRAPH", "0") == "1"
eng = engine_mod.TrainEngine(vae, B.CFG2, 128, B.SEQ_T, total_steps=B.TOTAL_STEPS, use_graph=use_graph, seed=10)
rng = np.random.default_rng(1000)
X, L, Y = B.synth_batch(rng, 128)
d = (torch.from_numpy(X).to(dev), torch.from_numpy(L).to(dev), torch.from_numpy(Y).to(dev))
for _ in range(6): eng.step_resident(*d)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): eng.step_resident(*d)
    torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", "step_trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# last step = kernels after the second-to-last clip_adam
adam = [i for i, e in enumerate(ev) if "clip_adam" in e["name"]]
step = ev[adam[-2] + 1: adam[-1] + 1]
t0 = step[0]["ts"]
print(f"step: {len(step)} kernels, span {(step[-1]['ts'] + step[-1]['dur'] - t0):.1f} us, summed {sum(e['dur'] for e in step):.1f} us")
end = t0
for e in step:
    gap = e["ts"] - end
    name = e["name"].split("(")[0].replace("void ", "").replace("dvae::", "").replace("(anonymous namespace)::", "")[:42]
    g = e["args"].get("grid", "")
    print(f"+{e['ts'] - t0:8.1f} dur {e['dur']:6.1f} stream {e['args'].get('stream', '?'):>3} {'GAP %5.1f' % gap if gap > 1.0 else '         '} {name} {g}")
    end = max(end, e["ts"] + e["dur"])
