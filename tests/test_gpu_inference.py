"""Encode-once / resample-R inference (scripts/evaluation/consistency.py:163-205) through `inference.ConsistencyEvaluator`:
the on-device length recount, graph replay == eager launches, and the second forward == the drop-in module on the sampled
sentences."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _params(**over):
    p = dict(bow_encoder=False, embedding_dim=32, hidden_dim=32, num_rnn_layers=2, encoder_dropout=0.0,
             decoder_dropout=0.0, bidirectional_encoder=True, latent_dims={"total": 8, "polarity": 1, "uncertainty": 1},
             adversarial_loss=False, mi_loss=False)
    p.update(over)
    return p


def _synthetic(B, T, V, gen):
    lengths = torch.randint(3, T + 1, (B,), generator=gen)
    lengths[0] = T
    X = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        n = int(lengths[b])
        X[b, 0], X[b, n - 1] = 2, 3
        X[b, 1:n - 1] = torch.randint(4, V, (n - 2,), generator=gen)
    return X, lengths


@pytest.mark.parametrize("B,T,stride_pad", [(1, 1, 0), (7, 9, 0), (33, 70, 3), (1024, 22, 0)])
def test_recount_lengths_matches_consistency_py(dvae, B, T, stride_pad):
    L = dvae._lib
    lib = L.load()
    gen = torch.Generator().manual_seed(B + T)
    big = torch.randint(0, 6, (B, T + stride_pad), generator=gen)          # ids 0..5: PAD = 0 and EOS = 3 are frequent
    if B > 2:
        big[1] = 3                                                           # a row of nothing but <EOS>
        big[2] = 5                                                           # a row with neither
    tok = big.cuda()[:, :T]                                                  # row stride T + stride_pad
    out = torch.full((B,), -7, device="cuda", dtype=torch.int64)
    for min_len in (0, 1):
        L.check(lib.dvae_recount_lengths(L.ptr(tok), tok.stride(0), tok.stride(1), B, T, 3, 0, min_len, L.ptr(out),
                                         L.stream_ptr()), "recount")
        x = big[:, :T]
        want = T - ((x == 3) | (x == 0)).sum(1)                              # consistency.py:186-190
        want = want.clamp(min=min_len)
        assert torch.equal(out.cpu(), want)


def test_resample_graph_equals_eager_and_hat_pass_equals_dropin_model(dvae):
    inf = importlib.import_module("disentanglement-vae_b200.inference")
    dvae.set_seed(10)
    V, B, T, R = 300, 24, 10, 3
    vae = dvae.build_vae(_params(), V, None, {"uncertainty": 1, "polarity": 1}, torch.device("cuda"), 2, 3)
    vae.eval()                                    # no dropout: the hat pass can be replayed through the module surface
    gen = torch.Generator().manual_seed(6)
    X, lengths = _synthetic(B, T, V, gen)
    outs = []
    for use_graph in (True, False):
        ev = inf.ConsistencyEvaluator(vae, B, T, use_graph=use_graph, seed=21)
        ctx = ev.encode_once(X, lengths)
        _, ctx_ref, _ = vae.encode(X.cuda(), lengths.cuda())
        assert torch.equal(ctx, ctx_ref)
        outs.append({k: v.clone() for k, v in ev.resample(R).items()})
    g, e = outs
    for k in g:
        assert torch.equal(g[k], e[k]), k          # same seed => same eps, same Gumbel noise, same tokens, same logits
    tok = g["token_predictions"]
    assert (tok[:, :, 0] == 2).all() and ((tok >= 0) & (tok < V)).all()
    assert not torch.equal(tok[0], tok[1])          # resamples differ
    want_len = (T - ((tok == 3) | (tok == 0)).sum(-1)).clamp(min=1)
    assert torch.equal(g["lengths_hat"], want_len)
    # first pass: z = mu + eps * exp(logvar) from the SAME context for every resample; logits from the fused heads
    P1 = vae.compute_latent_params(ctx_ref)
    mu = torch.cat([p.mu for p in P1.values()], 1)
    lv = torch.cat([p.logvar for p in P1.values()], 1)
    for r in range(R):
        eps_r = (g["z"][r] - mu) / lv.exp()
        assert 0.5 < eps_r.std().item() < 1.5      # fresh N(0,1) noise
    # second pass == module surface on (x_hat, lengths_hat): encode -> heads with the eps the evaluator drew
    r = R - 1
    _, ctx_hat, _ = vae.encode(tok[r], g["lengths_hat"][r])
    Ph = vae.compute_latent_params(ctx_hat)
    mu_h = torch.cat([p.mu for p in Ph.values()], 1)
    lv_h = torch.cat([p.logvar for p in Ph.values()], 1)
    eps_h = ev.plan.eps                              # eps of the last heads call (the hat pass of the last resample)
    assert torch.allclose(g["z_hat"][r], mu_h + eps_h * lv_h.exp(), rtol=1e-5, atol=1e-6)
    preds = ev.predictions(g["dsc_logits_hat"])
    assert set(preds) == {"uncertainty", "polarity"} and preds["polarity"].shape == (R, B)
    # discriminator logits of the hat pass = discriminators applied to z_hat
    off = 0
    for n, zs in zip(ev.d.space_names, ev.d.space_dims):
        if n in vae.discriminators:
            lin = vae.discriminators[n].linear
            want = g["z_hat"][r][:, off:off + zs] @ lin.weight.t() + lin.bias
            col = list(preds).index(n)
            assert torch.allclose(g["dsc_logits_hat"][r][:, col:col + 1], want, rtol=1e-4, atol=1e-5)
        off += zs


def test_resample_in_train_mode_reencode_each_draws_new_encoder_dropout(dvae):
    inf = importlib.import_module("disentanglement-vae_b200.inference")
    dvae.set_seed(3)
    V, B, T = 200, 16, 8
    vae = dvae.build_vae(_params(encoder_dropout=0.5, decoder_dropout=0.5), V, None, {"uncertainty": 1, "polarity": 1},
                         torch.device("cuda"), 2, 3)
    vae.train()                                     # consistency.py:151
    X, lengths = _synthetic(B, T, V, torch.Generator().manual_seed(1))
    ev = inf.ConsistencyEvaluator(vae, B, T, seed=4)
    ev.encode_once(X, lengths)
    c0 = ev.ctx0.clone()
    ev.resample(2)
    assert torch.equal(ev.ctx0, c0)                 # encode-once: the context is reused
    ev.resample(1, reencode_each=True)
    assert not torch.equal(ev.ctx0, c0)             # reference behaviour: a new dropout draw per forward
    with pytest.raises(dvae.DvaeError):
        inf.ConsistencyEvaluator(vae, B, T).resample(1)
