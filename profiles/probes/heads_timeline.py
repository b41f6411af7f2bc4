"""Phases of heads_fwd_kernel (CTA 0) inside a cfg-2 train step (eager launches): globaltimer marks via DVAE_HEADS_DBG."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import __graft_entry__ as ge
dvae = ge.build()
engine_mod = importlib.import_module("disentanglement-vae_b200.engine")
dev = torch.device("cuda", 0)
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
os.environ["DVAE_HEADS_DBG"] = hex(dbg.data_ptr())
B.select_workload("cfg2")
dvae.set_seed(10)
vae = dvae.build_vae(B.CFG2, B.VOCAB, None, B.LABELS, dev, B.SOS, B.EOS); vae.train()
eng = engine_mod.TrainEngine(vae, B.CFG2, 128, B.SEQ_T, total_steps=B.TOTAL_STEPS, use_graph=False, seed=10)
rng = np.random.default_rng(1000)
X, L, Y = B.synth_batch(rng, 128)
d = (torch.from_numpy(X).to(dev), torch.from_numpy(L).to(dev), torch.from_numpy(Y).to(dev))
for _ in range(6): eng.step_resident(*d)
torch.cuda.synchronize()
t = dbg.cpu().tolist()
names = ["kernel entry", "context rows staged", "context2params done", "reparameterisation + discriminators done", "z2hidden done",
         "per-CTA partials written", "exit (CTA 0)"]
for i, n in enumerate(names):
    print(f"  {n:42s} +{(t[i] - t[0]) / 1e3:6.2f} us")
