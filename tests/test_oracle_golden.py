"""Pin the oracle (`oracle/dvae_oracle.py`) against vectors produced by the unmodified reference.

The golden files were written by `tests/golden/make_golden.py`, which imports `/root/reference/vae`
and runs `build_vae` -> `forward` -> `compute_all_losses` -> `backward` -> clip -> `Adam.step`.
"""
import numpy as np
import pytest

from conftest import load_golden, golden_state_dict
from oracle import dvae_oracle as O


def _setup(name, dtype=np.float64):
    g = load_golden(name)
    sd = O.cast_state_dict(golden_state_dict(g), dtype)
    spec = O.ModelSpec(sd, [str(s) for s in g["space_names"]], int(g["sos"]), int(g["eos"]))
    eps = {n: g[f"eps.{n}"] for n in spec.space_names}
    labels = {str(n): g[f"Y.{n}"] for n in g["label_names"]}
    klw = {n: float(g[f"klw.{n}"]) for n in spec.space_names}
    fw = O.model_forward(sd, spec, g["inputs"], g["lengths"], eps, labels=labels, kl_weights=klw)
    return g, sd, spec, fw


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("case", ["tiny_uni", "tiny_bi", "tiny_eval_mc"])
def test_forward_matches_reference(case):
    g, sd, spec, fw = _setup(case)
    assert _rel(fw["context"], g["context"]) < 1e-5
    assert _rel(fw["enc_hn"], g["enc_hn"]) < 1e-5
    for n in spec.space_names:
        for f in ("z", "mu", "logvar"):
            assert _rel(fw["latent"][n][f], g[f"{f}.{n}"]) < 1e-5, (n, f)
        assert abs(fw["kls"][n] - float(g[f"kl.{n}"])) < 1e-5 * max(1.0, abs(float(g[f"kl.{n}"])))
    for n in spec.dsc_names:
        assert _rel(fw["dsc_logits"][n], g[f"dsc_logits.{n}"]) < 1e-5
        assert abs(fw["dsc_losses"][n] - float(g[f"dsc_loss.{n}"])) < 1e-5
        assert abs(fw["dsc_accs"][n] - float(g[f"dsc_acc.{n}"])) < 1e-6
    assert _rel(fw["dec_h0"], g["dec_h0"]) < 1e-5
    assert _rel(fw["dec_c0"], g["dec_c0"]) < 1e-5
    assert _rel(fw["decoder_logits"], g["decoder_logits"]) < 2e-5
    # token-level argmax identical (BASELINE.json north_star)
    assert np.array_equal(fw["decoder_logits"].argmax(-1), g["decoder_logits"].argmax(-1))
    # tf=1.0: token_predictions are the forced inputs (model.py:464-472)
    assert np.array_equal(g["token_predictions"][:, 1:], g["inputs"][:, 1:])
    for k in ("reconstruction", "total_weighted_kl", "total_kl", "total_dsc", "total"):
        want = float(g[f"loss.{k}"])
        got = {"reconstruction": fw["reconstruction_loss"], "total_weighted_kl": fw["total_weighted_kl"],
               "total_kl": fw["total_kl"], "total_dsc": fw["total_dsc_loss"], "total": fw["total_loss"]}[k]
        assert abs(got - want) <= 1e-5 * max(1.0, abs(want)), k


def test_position0_constant():
    # NLL at t=0 is log(e + V - 1) - 1 for every row whose first target is <SOS> (model.py:454)
    g, sd, spec, fw = _setup("tiny_uni")
    nll0 = fw["_bw"]["ce_cache"]["nll"][:, 0]
    assert np.allclose(nll0, np.log(np.e + spec.V - 1) - 1.0, rtol=1e-12)


@pytest.mark.parametrize("case", ["tiny_uni", "tiny_bi"])
def test_gradients_match_reference_autograd(case):
    g, sd, spec, fw = _setup(case)
    grads = O.model_backward(sd, spec, fw)
    checked = 0
    for k in sd:
        want = g[f"grad.{k}"]
        assert grads[k].shape == want.shape
        denom = max(np.abs(want).max(), 1e-8)
        assert np.abs(grads[k] - want).max() / denom < 2e-4, k      # fp32 autograd vs fp64 oracle
        checked += 1
    assert checked == len(sd)
    norm = np.sqrt(sum((grads[k] ** 2).sum() for k in sd))
    assert abs(norm - float(g["grad_norm"])) < 1e-4 * float(g["grad_norm"])


@pytest.mark.parametrize("case", ["tiny_uni", "tiny_bi"])
def test_clip_and_adam_match_reference(case):
    g, sd, spec, fw = _setup(case)
    grads = O.model_backward(sd, spec, fw)
    params = {k: v.copy() for k, v in sd.items()}
    m = {k: np.zeros_like(v) for k, v in sd.items()}
    v = {k: np.zeros_like(v_) for k, v_ in sd.items()}
    O.clip_and_adam(params, grads, m, v, step=1, lr=float(g["lr"]))
    for k in sd:
        want = g[f"sd_after.{k}"]
        # first Adam step moves every touched weight by ~lr; compare the update, not the weight
        upd_want = want.astype(np.float64) - g[f"sd.{k}"].astype(np.float64)
        upd_got = params[k] - sd[k]
        big = np.abs(g[f"grad.{k}"]) > 1e-6 * np.abs(g[f"grad.{k}"]).max()
        assert np.abs(upd_got - upd_want)[big].max() < 2e-2 * float(g["lr"]), k


def test_cyclic_schedule_known_answers():
    tab = load_golden("cyclic_kl")["table"]
    for step, total, want in tab:
        assert abs(O.cyclic_kl_weight(int(step), int(total)) - want) < 1e-12
    # SURVEY.md 8(a9) probed values, total=100
    for s, w in zip((0, 1, 5, 12, 13, 24, 25, 26), (0, .08, .4, .96, 1, 1, 0, .08)):
        assert abs(O.cyclic_kl_weight(s, 100) - w) < 1e-12


def test_float32_oracle_close_to_float64():
    g, sd, spec, fw64 = _setup("tiny_bi", np.float64)
    _, _, _, fw32 = _setup("tiny_bi", np.float32)
    assert abs(fw32["total_loss"] - fw64["total_loss"]) < 1e-5 * abs(fw64["total_loss"])


def test_oracle_auxiliary_objectives_match_reference_golden():
    """entropy loss of the adversaries, CLUB MI estimate (x beta 0.01) and CLUB learning loss, restated in numpy, against
    what the unmodified reference computed (tests/golden/tiny_adv_mi.npz, make_golden.py --adv-mi)."""
    g = load_golden("tiny_adv_mi")
    for n in [str(x) for x in g["adv_names"]]:
        loss, _ = O.entropy_loss(g[f"adv_logits.{n}"].astype(np.float64))
        assert abs(loss - float(g[f"adv_loss.{n}"])) < 1e-6, n
    total = 0.0
    for n in [str(x) for x in g["mi_names"]]:
        n1, n2 = n.split("-")
        W = {k[len(f"mi0.{n}."):]: g[k].astype(np.float64) for k in g if k.startswith(f"mi0.{n}.")}
        x, y = g[f"z.{n1}"].astype(np.float64), g[f"z.{n2}"].astype(np.float64)
        mu, _ = O.club_mlp(x, W, "p_mu")
        lv, _ = O.club_mlp(x, W, "p_logvar")
        mi = O.club_mi(mu, lv, y)[0] * 0.01
        assert abs(mi - float(g[f"mi_est.{n}"])) < 1e-7, n
        total += mi
        assert abs(O.club_nll(mu, lv, y)[0] - float(g[f"mi_learning_loss.{n}"])) < 1e-5, n
    assert abs(total - float(g["loss.total_mi"])) < 1e-7
    assert abs(sum(float(g[f"adv_loss.{n}"]) for n in g["adv_names"]) - float(g["loss.total_adv"])) < 1e-5


def test_oracle_bow_encoder_matches_reference_golden():
    """BOWEncoder (vae/model.py:13-49) restated in the oracle: loss, context and every gradient of the reference."""
    g = load_golden("tiny_bow")
    sd = O.cast_state_dict(golden_state_dict(g))
    spec = O.ModelSpec(sd, [str(s) for s in g["space_names"]], int(g["sos"]), int(g["eos"]))
    assert spec.bow and spec.C == spec.E
    fw = O.model_forward(sd, spec, g["inputs"], g["lengths"], {n: g[f"eps.{n}"] for n in spec.space_names},
                         labels={str(n): g[f"Y.{n}"] for n in g["label_names"]},
                         kl_weights={n: float(g[f"klw.{n}"]) for n in spec.space_names})
    assert np.abs(fw["context"] - g["context"]).max() < 1e-7
    assert abs(fw["total_loss"] - float(g["loss.total"])) < 1e-5 * abs(float(g["loss.total"]))
    grads = O.model_backward(sd, spec, fw)
    for k in sd:
        if f"grad.{k}" in g:
            assert np.abs(grads[k] - g[f"grad.{k}"]).max() <= 2e-4 * max(np.abs(g[f"grad.{k}"]).max(), 1e-8), k
