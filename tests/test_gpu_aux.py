"""Adversarial and mutual-information objectives (vae/model.py:219-258,323-355, vae/losses.py:10-74,
199-242, run.py:254-276): kernels vs the numpy oracle, and a full train step -- forward, compute_all_losses, backward,
clip, adversary steps, VAE Adam step, CLUB learning steps -- vs golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_state_dict
from oracle import dvae_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _rel(a, b):
    a = a.detach().cpu().numpy().astype(np.float64) if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("B,Oo", [(6, 1), (128, 1), (7, 3), (300, 5)])
def test_entropy_loss_kernel(dvae, B, Oo):
    L = dvae._lib
    lib = L.load()
    rng = np.random.default_rng(B + Oo)
    x = (rng.standard_normal((B, Oo)) * 3).astype(np.float32)
    x[0, 0] = 40.0                                         # saturated probability: exercises the clamp and its zero gradient
    xd = torch.from_numpy(x).cuda()
    loss = torch.zeros(1, device="cuda")
    d = torch.zeros(B, Oo, device="cuda")
    g = torch.tensor([0.7], device="cuda")
    st = L.stream_ptr()
    L.check(lib.dvae_entropy_loss(L.ptr(xd), B, Oo, L.ptr(loss), None, None, st), "fwd")
    L.check(lib.dvae_entropy_loss(L.ptr(xd), B, Oo, None, L.ptr(g), L.ptr(d), st), "bwd")
    want, dw = O.entropy_loss(x.astype(np.float64))
    assert abs(loss.item() - want) <= 1e-5 * max(abs(want), 1e-3)
    assert np.abs(d.cpu().numpy() - 0.7 * dw).max() <= 1e-5 * max(np.abs(dw).max(), 1e-6) + 1e-9


@pytest.mark.parametrize("B,D", [(6, 1), (128, 62), (50, 7)])
def test_club_kernels(dvae, B, D):
    L = dvae._lib
    lib = L.load()
    rng = np.random.default_rng(B * D)
    mu = rng.standard_normal((B, D)).astype(np.float32)
    lv = np.tanh(rng.standard_normal((B, D))).astype(np.float32)
    y = rng.standard_normal((B, D)).astype(np.float32)
    md, ld, yd = (torch.from_numpy(a).cuda() for a in (mu, lv, y))
    st = L.stream_ptr()
    out = torch.zeros(1, device="cuda")
    d_mu, d_lv, d_y = (torch.zeros(B, D, device="cuda") for _ in range(3))
    ws = torch.zeros(3 * D, device="cuda")
    g = torch.tensor([0.01], device="cuda")
    L.check(lib.dvae_club_mi(L.ptr(md), L.ptr(ld), L.ptr(yd), B, D, L.ptr(out), None, None, None, None, None, st), "mi")
    L.check(lib.dvae_club_mi(L.ptr(md), L.ptr(ld), L.ptr(yd), B, D, None, L.ptr(g), L.ptr(d_mu), L.ptr(d_lv), L.ptr(d_y), L.ptr(ws), st), "mi bwd")
    mi, a, b, c = O.club_mi(mu.astype(np.float64), lv.astype(np.float64), y.astype(np.float64))
    assert abs(out.item() - mi) <= 1e-5 * max(abs(mi), 1e-2)
    assert _rel(d_mu, 0.01 * a) < 1e-4 and _rel(d_lv, 0.01 * b) < 1e-4 and _rel(d_y, 0.01 * c) < 1e-4
    L.check(lib.dvae_club_nll(L.ptr(md), L.ptr(ld), L.ptr(yd), B, D, L.ptr(out), None, None, None, st), "nll")
    L.check(lib.dvae_club_nll(L.ptr(md), L.ptr(ld), L.ptr(yd), B, D, None, None, L.ptr(d_mu), L.ptr(d_lv), st), "nll bwd")
    nll, a, b = O.club_nll(mu.astype(np.float64), lv.astype(np.float64), y.astype(np.float64))
    assert abs(out.item() - nll) <= 1e-5 * abs(nll)
    assert _rel(d_mu, a) < 1e-5 and _rel(d_lv, b) < 1e-5
    # activation derivatives
    yv, gv = torch.from_numpy(np.tanh(mu)).cuda(), torch.from_numpy(y).cuda()
    dd = torch.zeros(B, D, device="cuda")
    L.check(lib.dvae_act_bwd(L.ptr(yv), L.ptr(gv), L.ptr(dd), B * D, 1, st), "tanh bwd")
    assert _rel(dd, y * (1 - np.tanh(mu) ** 2)) < 1e-6
    rl = md.clone()
    L.check(lib.dvae_relu(L.ptr(rl), B * D, st), "relu")
    assert np.array_equal(rl.cpu().numpy(), np.maximum(mu, 0))
    L.check(lib.dvae_act_bwd(L.ptr(rl), L.ptr(gv), L.ptr(dd), B * D, 2, st), "relu bwd")
    assert np.array_equal(dd.cpu().numpy(), np.where(mu > 0, y, 0).astype(np.float32))


def test_adversarial_mi_train_step_matches_reference_golden(dvae):
    g = load_golden("tiny_adv_mi")
    sd = golden_state_dict(g)
    names = [str(s) for s in g["space_names"]]
    dims = [int(x) for x in g["space_dims"]]
    label_dims = {str(n): int(d) for n, d in zip(g["label_names"], g["label_dims"])}
    lat = {"total": sum(dims)}
    for n, zs in zip(names, dims):
        if n != "content":
            lat[n] = zs
    p = dict(bow_encoder=False, embedding_dim=sd["encoder.embedding.weight"].shape[1],
             hidden_dim=sd["decoder.recurrent.weight_hh_l0"].shape[1], num_rnn_layers=2, encoder_dropout=0.0,
             decoder_dropout=0.0, bidirectional_encoder=True, latent_dims=lat, adversarial_loss=True, mi_loss=True)
    dev = torch.device("cuda")
    vae = dvae.build_vae(p, int(g["V"]), None, label_dims, dev, int(g["sos"]), int(g["eos"]))
    vae.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    for n, est in vae.mi_estimators.items():
        est.load_state_dict({k[len(f"mi0.{n}."):]: torch.from_numpy(g[k]) for k in g if k.startswith(f"mi0.{n}.")})
    assert list(vae.adversaries.keys()) == [str(x) for x in g["adv_names"]]
    assert list(vae.mi_estimators.keys()) == [str(x) for x in g["mi_names"]]
    vae.train()
    X = torch.from_numpy(g["inputs"]).to(dev)
    lengths = torch.from_numpy(g["lengths"]).to(dev)
    eps = torch.from_numpy(np.concatenate([g[f"eps.{n}"] for n in names], axis=1)).to(dev)
    Y = {n: torch.from_numpy(g[f"Y.{n}"]) for n in label_dims}
    klw = {n: float(g[f"klw.{n}"]) for n in names}
    klw["default"] = klw["content"]
    opt = torch.optim.Adam(vae.trainable_parameters(), lr=float(g["lr"]))
    out = vae(X, lengths, teacher_forcing_prob=1.0, eps=eps)
    for n in vae.adversaries:
        assert _rel(out["adv_logits"][n], g[f"adv_logits.{n}"]) < 1e-5
    total, Ls = dvae.losses.compute_all_losses(vae, out, X, Y, lengths, klw, mi_loss_weight=0.01)
    assert abs(total.item() - float(g["loss.total"])) <= 1e-5 * abs(float(g["loss.total"]))
    for n in vae.adversaries:
        assert abs(Ls["idv_adv_losses"][n] - float(g[f"adv_loss.{n}"])) < 1e-5
        assert abs(Ls["idv_adv_dsc_losses"][n].item() - float(g[f"adv_dsc_loss.{n}"])) < 1e-5
        assert abs(Ls["idv_adv_dsc_accs"][n] - float(g[f"adv_dsc_acc.{n}"])) < 1e-6
    for n in vae.mi_estimators:
        assert abs(Ls["idv_mi_estimates"][n] - float(g[f"mi_est.{n}"])) < 1e-6
    # run.py:254-262
    total.backward(retain_graph=True)
    for k, prm in vae.named_parameters():
        assert _rel(prm.grad, g[f"grad.{k}"]) < 1e-3, k            # adversaries.*: the entropy-term gradients
    norm = torch.nn.utils.clip_grad_norm_(vae.trainable_parameters(), 5.0)
    assert abs(norm.item() - float(g["grad_norm"])) <= 1e-4 * float(g["grad_norm"])
    for n, dl in Ls["idv_adv_dsc_losses"].items():
        adv = vae.adversaries[n]
        dl.backward(retain_graph=True)
        for k, prm in adv.named_parameters():
            assert _rel(prm.grad, g[f"adv_grad_at_step.{n}.{k}"]) < 1e-3, (n, k)
        adv.optimizer.step()
        adv.optimizer.zero_grad()
    opt.step()
    opt.zero_grad()
    # run.py:264-276
    for n, est in vae.mi_estimators.items():
        n1, n2 = n.split("-")
        z1, z2 = out["latent_params"][n1].z.detach(), out["latent_params"][n2].z.detach()
        ll = est.learning_loss(z1, z2)
        assert abs(ll.item() - float(g[f"mi_learning_loss.{n}"])) <= 1e-5 * abs(float(g[f"mi_learning_loss.{n}"]))
        est.optimizer.zero_grad()
        ll.backward()
        for k, prm in est.named_parameters():
            assert _rel(prm.grad, g[f"mi_grad.{n}.{k}"]) < 1e-3, (n, k)
        torch.nn.utils.clip_grad_norm_(est.parameters(), 1.0)
        est.optimizer.step()
        for k, v in est.state_dict().items():
            assert np.abs(v.cpu().numpy() - g[f"mi_after.{n}.{k}"]).max() < 2e-6, (n, k)
    for k, v in vae.state_dict().items():
        assert np.abs(v.cpu().numpy() - g[f"sd_after.{k}"]).max() < 2e-6, k
    # optimizer_step as the reference's method (model.py:239-245) is the same sequence
    assert callable(vae.adversaries[list(vae.adversaries)[0]].optimizer_step)
