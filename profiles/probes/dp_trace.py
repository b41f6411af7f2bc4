"""Kernel timeline of one data-parallel cfg-2 train step on rank 0 (torchrun, graphs + bucketed all-reduce), from CUPTI via
torch.profiler.  Not a bench: tracing adds overhead."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench as B
import __graft_entry__ as ge
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
dvae = ge.build()
engine_mod = importlib.import_module("disentanglement-vae_b200.engine")
B.select_workload("cfg2")
dvae.set_seed(10)
vae = dvae.build_vae(B.CFG2, B.VOCAB, None, B.LABELS, dev, B.SOS, B.EOS); vae.train()
eng = engine_mod.TrainEngine(vae, B.CFG2, 128, B.SEQ_T, total_steps=B.TOTAL_STEPS, use_graph=os.environ.get("TRACE_GRAPH", "1") == "1",
                             seed=10, process_group=dist.group.WORLD)
rng = np.random.default_rng(1000 + rank)
X, L, Y = B.synth_batch(rng, 128)
d = (torch.from_numpy(X).to(dev), torch.from_numpy(L).to(dev), torch.from_numpy(Y).to(dev))
for _ in range(8): eng.step_resident(*d)
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4): eng.step_resident(*d)
    torch.cuda.synchronize()
if rank == 0:
    path = os.path.join(ROOT, "gpurun_out", "dp_trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    adam = [i for i, e in enumerate(ev) if "clip_adam" in e["name"]]
    step = ev[adam[-2] + 1: adam[-1] + 1]
    t0 = step[0]["ts"]
    print("periods between optimizer kernels (us):", [round(ev[adam[i + 1]]["ts"] - ev[adam[i]]["ts"], 1) for i in range(len(adam) - 1)])
    print(f"step: {len(step)} kernels, span {(step[-1]['ts'] + step[-1]['dur'] - t0):.1f} us, summed {sum(e['dur'] for e in step):.1f} us")
    end = t0
    for e in step:
        gap = e["ts"] - end
        name = e["name"].split("(")[0].replace("void ", "").replace("dvae::", "").replace("(anonymous namespace)::", "")[:42]
        g = e["args"].get("grid", "")
        print(f"+{e['ts'] - t0:8.1f} dur {e['dur']:6.1f} stream {e['args'].get('stream', '?'):>3} {'GAP %5.1f' % gap if gap > 1.0 else '         '} {name} {g}")
        end = max(end, e["ts"] + e["dur"])
    os.remove(path)
dist.barrier()
dist.destroy_process_group()
