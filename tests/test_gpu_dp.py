"""Hardware data-parallel parity (VERDICT r1 "next 1b"): N ranks x (B/N) rows through the NCCL
TrainEngine (CUDA graphs split around the bucketed all-reduce, communication stream overlap) must take the same
optimisation steps as ONE engine on the concatenated batch.  Spawns `torch.distributed.run`; skipped with < 2 GPUs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
def test_n_rank_nccl_engine_equals_single_rank_on_the_global_batch(tmp_path, world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / f"dp{world}.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py"), str(out)]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = np.load(out)
    assert got["replicas_identical"][0] == 1.0, "ranks diverged: parameters differ after the all-reduced steps"
    import dp_worker
    vae, eng, losses = dp_worker.run(1, 0, torch.device("cuda", 0))
    want = vae._flat.detach().cpu().numpy()
    assert int(got["adam_step"]) == eng.adam_step == dp_worker.STEPS
    for a, b in zip(got["losses"], losses):
        assert abs(a - b) <= 1e-5 * abs(b), (got["losses"], losses)
    # post-Adam weights: SURVEY 4(iv) asks 2e-3; Adam turns noise-level gradients into +-lr moves, so the bound is on the
    # update relative to the largest update, over entries that moved at all
    import importlib
    dvae = importlib.import_module("disentanglement-vae_b200")
    dvae.set_seed(10)
    w0 = dvae.build_vae(dp_worker.make_cfg(), dp_worker.V, None, {"uncertainty": 1, "polarity": 1}, torch.device("cuda", 0), 2, 3)
    w0 = w0._flat.detach().cpu().numpy()[:want.size]
    upd_want, upd_got = want - w0, got["flat"][:want.size] - w0
    scale = np.abs(upd_want).max()
    frac_bad = (np.abs(upd_got - upd_want) > 2e-3 * scale).mean()
    assert frac_bad < 2e-3, frac_bad
    assert np.abs(upd_got - upd_want).mean() < 1e-3 * np.abs(upd_want).mean()
