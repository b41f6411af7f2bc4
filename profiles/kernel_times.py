"""Per-kernel device time of the eager (non-graph) train step via torch.profiler (CUPTI), warm caches.
Analysis aid only -- never a bench value.   python profiles/kernel_times.py [steps]"""
import collections
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import __graft_entry__ as ge  # noqa: E402
from importlib import import_module  # noqa: E402

dvae = ge.build()
engine_mod = import_module("disentanglement-vae_b200.engine")
dev = torch.device("cuda")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
uniform_lengths = bench.select_workload(workload)
dvae.set_seed(10)
vae = dvae.build_vae(bench.CFG2, bench.VOCAB, None, bench.LABELS, dev, bench.SOS, bench.EOS)
vae.train()
eng = engine_mod.TrainEngine(vae, bench.CFG2, 128, bench.SEQ_T, total_steps=bench.TOTAL_STEPS, use_graph=False)
rng = np.random.default_rng(1000)
X, L, Y = bench.synth_batch(rng, 128, uniform_lengths=uniform_lengths)
batch = (torch.from_numpy(X).to(dev), torch.from_numpy(L).to(dev), torch.from_numpy(Y).to(dev))
for _ in range(3):
    eng.step_resident(*batch)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        eng.step_resident(*batch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
by_grid = os.environ.get("KT_BY_GRID", "0") == "1"      # KT_BY_GRID=1: split each kernel name by launch grid (chrome trace)
if by_grid:
    import json
    path = os.path.join(ROOT, "gpurun_out", "kt_trace.json")
    prof.export_chrome_trace(path)
    for e in json.load(open(path))["traceEvents"]:
        if e.get("cat") == "kernel":
            nm = e["name"].split("(")[0].replace("void ", "").replace("dvae::", "").replace("(anonymous namespace)::", "")[:60]
            a = agg[f"{nm} {e['args'].get('grid', '')}"]
            a[0] += 1
            a[1] += e["dur"]
    os.remove(path)
for e in ([] if by_grid else prof.events()):
    if e.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[e.name[:100]]
        a[0] += 1
        a[1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"# {workload}: {steps} eager steps, {tot / steps:.1f} us kernel time per step")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:(60 if by_grid else 28)]:
    print(f"{t / steps:10.1f} us {100 * t / tot:5.1f}% x{c // steps:4d}  {k}")
