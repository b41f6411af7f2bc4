"""TrainEngine beyond the headline path: asynchronous step_resident (pinned-staging ring), teacher forcing < 1 with
device-side coins, the adversarial / MI step (run.py:254-276) against the reference golden vectors, and the optimizer
state in the reference's checkpoint format."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from conftest import load_golden, golden_state_dict  # noqa: E402
from oracle import dvae_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dvae():
    import __graft_entry__ as ge
    return ge.build()


@pytest.fixture(scope="module")
def engine_mod(dvae):
    return importlib.import_module("disentanglement-vae_b200.engine")


def _cfg(**over):
    c = dict(bow_encoder=False, embedding_dim=64, hidden_dim=64, num_rnn_layers=2, encoder_dropout=0.5, decoder_dropout=0.5,
             bidirectional_encoder=True, latent_dims={"total": 16, "polarity": 1, "uncertainty": 1}, adversarial_loss=False,
             mi_loss=False, learn_rate=3e-3, lambdas={"default": "cyclic", "polarity": 0.005, "uncertainty": 0.005},
             random_seed=10, teacher_forcing_prob=1.0)
    c.update(over)
    return c


def _batch(gen, B, T, V, names=("uncertainty", "polarity")):
    lengths = torch.randint(3, T + 1, (B,), generator=gen)
    lengths[0] = T
    X = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        n = int(lengths[b])
        X[b, 0], X[b, n - 1] = 2, 3
        X[b, 1:n - 1] = torch.randint(4, V, (n - 2,), generator=gen)
    Y = {k: (torch.rand(B, 1, generator=gen) < 0.3).float() for k in names}
    return X, lengths, Y


def _rel(a, b):
    a = a.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(b) else np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_step_resident_back_to_back_equals_step_host(dvae, engine_mod):
    """the per-step scalar block (Adam step, KL weight, Philox seed) goes through pinned memory with
    asynchronous copies.  Eight un-synchronised step_resident calls must see eight DIFFERENT blocks: same losses and
    the same weights as the synchronous step_host path under the same seed (dropout 0.5, so a repeated seed shows)."""
    V, B, T = 500, 24, 9
    dev = torch.device("cuda")
    runs = []
    for mode in ("host", "resident"):
        dvae.set_seed(10)
        vae = dvae.build_vae(_cfg(), V, None, {"uncertainty": 1, "polarity": 1}, dev, 2, 3)
        vae.train()
        w0 = vae._flat.detach().clone()
        eng = engine_mod.TrainEngine(vae, _cfg(), B, T, total_steps=40, use_graph=True, seed=5)
        gen = torch.Generator().manual_seed(1)
        batches = [_batch(gen, B, T, V) for _ in range(8)]
        outs = []
        if mode == "host":
            for X, L, Y in batches:
                outs.append(eng.step_host(X, L, Y)["total_loss"])
        else:
            dbat = [(X.to(dev), L.to(dev), torch.stack([Y[n].reshape(-1) for n in eng.label_names]).to(dev)) for X, L, Y in batches]
            torch.cuda.synchronize()
            blocks = [eng.step_resident(*b).clone() for b in dbat]        # no sync between the steps
            torch.cuda.synchronize()
            outs = [eng.losses_from(b.cpu())["total_loss"] for b in blocks]
        runs.append((outs, vae._flat.detach().clone(), eng.adam_step, w0))
    (la, wa, sa, w0), (lb, wb, sb, _) = runs
    assert sa == sb == 8
    assert len(set(round(x, 3) for x in lb)) == 8
    for a, b in zip(la, lb):
        assert abs(a - b) <= 1e-4 * abs(a), (la, lb)
    # weights: the two runs differ only by the summation order of atomics (Adam turns a sign flip of a noise-level
    # gradient into a +-lr move of that entry, so compare in the mean, against the size of the 8-step update)
    assert (wb - wa).abs().mean().item() < 0.02 * (wa - w0).abs().mean().item()


@pytest.mark.parametrize("use_graph", [False, True])
def test_engine_teacher_forcing_half_matches_oracle_on_sampled_tokens(dvae, engine_mod, use_graph):
    """params["teacher_forcing_prob"] = 0.5 (every shipped config): coins on the device, sampled inputs; the oracle run
    teacher-forced on the tokens the engine fed its decoder must give the same loss and the same Adam update."""
    V, B, T, lr, total = 400, 20, 10, 3e-3, 30
    cfg = _cfg(encoder_dropout=0.0, decoder_dropout=0.0, teacher_forcing_prob=0.5, learn_rate=lr)
    dev = torch.device("cuda")
    dvae.set_seed(10)
    vae = dvae.build_vae(cfg, V, None, {"uncertainty": 1, "polarity": 1}, dev, 2, 3)
    vae.train()
    eng = engine_mod.TrainEngine(vae, cfg, B, T, total_steps=total, use_graph=use_graph, seed=3)
    assert eng.sampled
    gen = torch.Generator().manual_seed(17)
    seen_sampled = 0
    for step in range(3):
        sd = O.cast_state_dict({k: v.detach().cpu().numpy() for k, v in vae.state_dict().items()})
        Mv, Vv = vae.grad_views(eng.m), vae.grad_views(eng.v)
        m = {k: Mv[k].detach().cpu().numpy().astype(np.float64) for k in sd}
        v2 = {k: Vv[k].detach().cpu().numpy().astype(np.float64) for k in sd}
        spec = O.ModelSpec(sd, list(vae.context2params.keys()), 2, 3)
        X, L, Y = _batch(gen, B, T, V)
        got = eng.step_host(X, L, Y)
        coins = eng.coins.cpu().numpy()
        preds = eng.preds.cpu().numpy()
        assert (preds[:, 0] == 2).all()
        for i in range(1, T):
            if coins[i - 1]:
                assert np.array_equal(preds[:, i], X[:, i].numpy())
            else:
                seen_sampled += 1
                assert (preds[:, i] != X[:, i].numpy()).mean() > 0.5
        eps = eng.plan.eps.detach().cpu().numpy()
        eps_d, off = {}, 0
        for n, zs in zip(spec.space_names, spec.space_dims):
            eps_d[n] = eps[:, off:off + zs]
            off += zs
        klw = {"default": O.cyclic_kl_weight(step, total), "polarity": 0.005, "uncertainty": 0.005}
        fw = O.model_forward(sd, spec, X.numpy(), L.numpy(), eps_d, labels={k: y.numpy() for k, y in Y.items()},
                             kl_weights=klw, dec_inputs=preds[:, :T - 1])
        grads = O.model_backward(sd, spec, fw)
        assert abs(got["total_loss"] - fw["total_loss"]) <= 1e-5 * abs(fw["total_loss"])
        before = {k: a.copy() for k, a in sd.items()}
        O.clip_and_adam(sd, grads, m, v2, step + 1, lr)
        now = {k: t.detach().cpu().numpy().astype(np.float64) for k, t in vae.state_dict().items()}
        for k in sd:
            sig = np.abs(grads[k]) > 1e-4 * np.abs(grads[k]).max()
            if sig.any():
                want, have = sd[k] - before[k], now[k] - before[k]
                assert np.abs(want - have)[sig].max() / max(np.abs(want[sig]).max(), 1e-30) < 2e-3, (step, k)
    assert seen_sampled > 0


def test_engine_rejects_cyclic_without_total_steps(dvae, engine_mod):
    vae = dvae.build_vae(_cfg(), 50, None, {"uncertainty": 1, "polarity": 1}, torch.device("cuda"), 2, 3)
    with pytest.raises(ValueError):
        engine_mod.TrainEngine(vae, _cfg(), 4, 5, total_steps=None)


def test_engine_optimizer_state_round_trips_with_torch_adam_and_checkpoints(dvae, engine_mod, tmp_path):
    """run.py:624-630 / vae/utils.py:147-175: a checkpoint written from the engine resumes (a) a torch.optim.Adam as the
    reference builds it and (b) another engine, which then takes the same next step."""
    V, B, T = 300, 12, 8
    cfg = _cfg(encoder_dropout=0.0, decoder_dropout=0.0)
    dev = torch.device("cuda")
    dvae.set_seed(10)
    vae = dvae.build_vae(cfg, V, None, {"uncertainty": 1, "polarity": 1}, dev, 2, 3)
    vae.train()
    eng = engine_mod.TrainEngine(vae, cfg, B, T, total_steps=20, use_graph=False, seed=2)
    gen = torch.Generator().manual_seed(4)
    for _ in range(2):
        eng.step_host(*_batch(gen, B, T, V))
    path = eng.save_checkpoint(str(tmp_path), 0)
    ck = torch.load(path)
    assert set(ck) == {"model_state_dict", "optimizer_state_dict", "epoch"}
    # (a) the reference's optimizer object accepts it
    opt = torch.optim.Adam(vae.trainable_parameters(), lr=1.0)
    opt.load_state_dict(ck["optimizer_state_dict"])
    Mv = vae.grad_views(eng.m)
    names = [n for n, _ in eng._trainable()]
    for n, p in zip(names, vae.trainable_parameters()):
        st = opt.state[p]
        assert float(st["step"]) == 2.0 and torch.equal(st["exp_avg"].to(dev), Mv[n])
    assert opt.param_groups[0]["lr"] == cfg["learn_rate"]
    # (b) a fresh engine resumes and reproduces the third step
    X, L, Y = _batch(gen, B, T, V)
    eps = torch.randn(B, 16, generator=gen).to(dev)
    eng.fixed_eps = eps
    want = eng.step_host(X, L, Y)
    w_want = vae._flat.detach().clone()
    dvae.set_seed(99)
    vae2 = dvae.build_vae(cfg, V, None, {"uncertainty": 1, "polarity": 1}, dev, 2, 3)
    vae2.train()
    eng2 = engine_mod.TrainEngine(vae2, cfg, B, T, total_steps=20, use_graph=False, seed=2)
    next_epoch, fname = eng2.load_latest_checkpoint(str(tmp_path), steps_per_epoch=2)
    assert (next_epoch, fname) == (1, "model_0.pt") and eng2.adam_step == 2 and eng2.step_idx == 2
    eng2.fixed_eps = eps
    got = eng2.step_host(X, L, Y)
    assert abs(got["total_loss"] - want["total_loss"]) <= 1e-6 * abs(want["total_loss"])
    assert (vae2._flat - w_want).abs().mean().item() < 1e-3 * cfg["learn_rate"]


def test_engine_adversarial_mi_step_matches_reference_golden(dvae, engine_mod):
    """The whole of run.py:217-276 through TrainEngine (adversaries + CLUB estimators) vs the unmodified reference:
    losses, VAE weights after clip + Adam, adversary and estimator weights after their own optimizers."""
    g = load_golden("tiny_adv_mi")
    sd = golden_state_dict(g)
    names = [str(s) for s in g["space_names"]]
    dims = [int(x) for x in g["space_dims"]]
    label_dims = {str(n): int(d) for n, d in zip(g["label_names"], g["label_dims"])}
    lat = {"total": sum(dims)}
    for n, zs in zip(names, dims):
        if n != "content":
            lat[n] = zs
    klw = {n: float(g[f"klw.{n}"]) for n in names}
    lambdas = {"default": klw["content"], **{n: klw[n] for n in names if n != "content"}}
    p = dict(bow_encoder=False, embedding_dim=sd["encoder.embedding.weight"].shape[1],
             hidden_dim=sd["decoder.recurrent.weight_hh_l0"].shape[1], num_rnn_layers=2, encoder_dropout=0.0,
             decoder_dropout=0.0, bidirectional_encoder=True, latent_dims=lat, adversarial_loss=True, mi_loss=True,
             learn_rate=float(g["lr"]), lambdas=lambdas, teacher_forcing_prob=1.0, random_seed=10)
    dev = torch.device("cuda")
    vae = dvae.build_vae(p, int(g["V"]), None, label_dims, dev, int(g["sos"]), int(g["eos"]))
    vae.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    for n, est in vae.mi_estimators.items():
        est.load_state_dict({k[len(f"mi0.{n}."):]: torch.from_numpy(g[k]) for k in g if k.startswith(f"mi0.{n}.")})
    vae.train()
    X, lengths = torch.from_numpy(g["inputs"]), torch.from_numpy(g["lengths"])
    B, T = X.shape
    eng = engine_mod.TrainEngine(vae, p, B, T, total_steps=10, seed=1)
    assert eng.aux and not eng.use_graph
    eng.fixed_eps = torch.from_numpy(np.concatenate([g[f"eps.{n}"] for n in names], axis=1)).to(dev)
    Y = {n: torch.from_numpy(g[f"Y.{n}"]) for n in label_dims}
    got = eng.step_host(X, lengths, Y)
    assert abs(got["total_loss"] - float(g["loss.total"])) <= 1e-5 * abs(float(g["loss.total"]))
    for n in vae.adversaries:
        assert abs(got["idv_adv_losses"][n] - float(g[f"adv_loss.{n}"])) < 1e-5
    for n in vae.mi_estimators:
        assert abs(got["idv_mi_estimates"][n] - float(g[f"mi_est.{n}"])) < 1e-6
    for k, v in vae.state_dict().items():          # VAE weights after clip + Adam, adversaries after their own Adam
        assert np.abs(v.cpu().numpy() - g[f"sd_after.{k}"]).max() < 2e-6, k
    for n, est in vae.mi_estimators.items():
        for k, v in est.state_dict().items():
            assert np.abs(v.cpu().numpy() - g[f"mi_after.{n}.{k}"]).max() < 2e-6, (n, k)


def test_pipelined_train_steps_equal_synchronous_steps(dvae, engine_mod):
    """TrainEngine.train_steps (batch k + 1 staged while step k runs, losses read two steps behind) takes exactly the steps
    of one step_host call per batch: same per-step losses under the same seed, every batch consumed, in order."""
    V, B, T = 500, 24, 9
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(3)
    batches = [_batch(gen, B, T, V) for _ in range(7)]
    runs = []
    for mode in ("sync", "pipe"):
        dvae.set_seed(10)
        vae = dvae.build_vae(_cfg(), V, None, {"uncertainty": 1, "polarity": 1}, dev, 2, 3)
        vae.train()
        eng = engine_mod.TrainEngine(vae, _cfg(), B, T, total_steps=40, use_graph=True, seed=5)
        if mode == "sync":
            runs.append([eng.step_host(*b)["total_loss"] for b in batches])
        else:
            runs.append([L["total_loss"] for L in eng.train_steps(iter(batches))])
        assert eng.adam_step == 7 and eng.step_idx == 7
    a, b = runs
    assert len(a) == len(b) == 7 and len(set(round(x, 3) for x in b)) == 7
    for x, y in zip(a, b):
        assert abs(x - y) <= 1e-4 * abs(x), (a, b)
