"""One eager train step (no CUDA graph, so every kernel is a separate launch) between cudaProfilerStart / Stop:

    python profiles/probes/ncu_step.py [cfg2|cfg4|cfg1] &&
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r2_step \
        python profiles/probes/ncu_step.py [cfg2|cfg4|cfg1]

Every launch of the step is captured once (~100 kernels at cfg 2); profiles/summarize_ncu.py turns the report into text."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import __graft_entry__ as ge
dvae = ge.build()
engine_mod = importlib.import_module("disentanglement-vae_b200.engine")
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
ul = B.select_workload(wl)
dev = torch.device("cuda", 0)
dvae.set_seed(10)
vae = dvae.build_vae(B.CFG2, B.VOCAB, None, B.LABELS, dev, B.SOS, B.EOS); vae.train()
eng = engine_mod.TrainEngine(vae, B.CFG2, B.BATCH, B.SEQ_T, total_steps=B.TOTAL_STEPS, use_graph=False, seed=10)
rng = np.random.default_rng(1000)
X, L, Y = B.synth_batch(rng, B.BATCH, uniform_lengths=ul)
d = (torch.from_numpy(X).to(dev), torch.from_numpy(L).to(dev), torch.from_numpy(Y).to(dev))
for _ in range(3):
    eng.step_resident(*d)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step_resident(*d)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", eng.losses_from(eng.plan.out.cpu())["total_loss"])
