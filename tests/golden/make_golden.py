"""Generate golden vectors by running the UNMODIFIED reference (`/root/reference/vae`) on CPU.

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py

Writes `tests/golden/<case>.npz` (small, committed).  Each file holds the reference's own
`state_dict`, the seeded synthetic batch, the replayed reparameterisation noise, and what the
reference computed from them: `decoder_logits`, latent params, every loss term
(`run.py:128-163` `compute_all_losses`), every gradient after `total_loss.backward()`, the
clip norm (`run.py:255`) and the parameters after one `Adam.step()` (`run.py:261`).

Parity-mode settings: dropout 0.0, teacher_forcing_prob 1.0, train mode, eps
replayed in the reference's order (2 draws per space in train mode, 2nd used; `model.py:391-395`).
"""
import os
import sys
import random

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402

ref_model, ref_losses, ref_utils = load_reference()

PAD, UNK, SOS, EOS = 0, 1, 2, 3


def make_params(**over):
    p = {"name": "golden", "random_seed": 10, "data_dir": "", "combined_dataset": False,
         "dataset_minibatch_ratios": {}, "checkpoint_dir": "", "glove_path": "",
         "num_train_examples": -1, "lowercase": True, "reverse_input": False,
         "embedding_dim": 12, "hidden_dim": 16, "num_rnn_layers": 2,
         "bidirectional_encoder": False, "bow_encoder": False,
         "latent_dims": {"total": 6, "polarity": 1}, "epochs": 3, "batch_size": 5,
         "learn_rate": 5e-3, "encoder_dropout": 0.0, "decoder_dropout": 0.0,
         "teacher_forcing_prob": 1.0, "lambdas": {"default": 0.01, "polarity": 0.005},
         "adversarial_loss": False, "mi_loss": False, "train": True, "validate": False,
         "test": False}
    p.update(over)
    ref_utils.validate_params(p)
    return p


def make_batch(gen, B, T, V, label_dims, full_row=True):
    lengths = torch.randint(3, T + 1, (B,), generator=gen)
    if full_row:
        lengths[gen.initial_seed() % B] = T
    X = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        n = int(lengths[b])
        X[b, 0] = SOS
        X[b, 1:n - 1] = torch.randint(4, V, (n - 2,), generator=gen)
        X[b, n - 1] = EOS
    Y = {}
    for name, dim in label_dims.items():
        if dim == 1:
            Y[name] = (torch.rand(B, 1, generator=gen) < 0.4).float()
        else:
            Y[name] = torch.randint(0, dim, (B, 1), generator=gen)
    return X, lengths, Y


def run_case(name, params, V, label_dims, B, T, kl_weights, seed=10, eps_seed=1234, mode="train"):
    ref_utils.set_seed(seed)
    vae = ref_model.build_vae(params, V, None, label_dims, torch.device("cpu"), SOS, EOS)
    gen = torch.Generator().manual_seed(seed + 1)
    X, lengths, Y = make_batch(gen, B, T, V, label_dims)
    out = {"B": B, "T": T, "V": V, "sos": SOS, "eos": EOS, "inputs": X.numpy(),
           "lengths": lengths.numpy(), "lr": params["learn_rate"]}
    for k, v in Y.items():
        out[f"Y.{k}"] = v.numpy()
    for k, v in vae.state_dict().items():
        out[f"sd.{k}"] = v.detach().numpy().copy()
    space_names = list(vae.context2params.keys())
    out["space_names"] = np.array(space_names)
    out["space_dims"] = np.array([vae.context2params[n].out_features // 2 for n in space_names])
    out["label_names"] = np.array(list(label_dims.keys()))
    out["label_dims"] = np.array(list(label_dims.values()))
    for k in space_names:
        out[f"klw.{k}"] = np.float64(kl_weights.get(k, kl_weights["default"]))

    vae.train(mode == "train")
    # replay the reference's noise order (model.py:387-396)
    torch.manual_seed(eps_seed)
    for n in space_names:
        zs = vae.context2params[n].out_features // 2
        if mode == "train":
            torch.randn(B, zs)
        out[f"eps.{n}"] = torch.randn(B, zs).numpy()
    torch.manual_seed(eps_seed)
    random.seed(seed)
    opt = torch.optim.Adam(vae.trainable_parameters(), lr=params["learn_rate"])
    output = vae(X, lengths, teacher_forcing_prob=params["teacher_forcing_prob"])
    for n in space_names:
        P = output["latent_params"][n]
        z_replay = P.mu + torch.from_numpy(out[f"eps.{n}"]) * torch.exp(P.logvar)
        assert torch.equal(z_replay, P.z), f"eps replay failed for {n}"
        out[f"z.{n}"] = P.z.detach().numpy()
        out[f"mu.{n}"] = P.mu.detach().numpy()
        out[f"logvar.{n}"] = P.logvar.detach().numpy()
    for n, l in output["dsc_logits"].items():
        out[f"dsc_logits.{n}"] = l.detach().numpy()
    out["decoder_logits"] = output["decoder_logits"].detach().numpy()
    out["token_predictions"] = output["token_predictions"].numpy()
    if params["bow_encoder"] is True:
        out["context"] = vae.encoder(X).detach().numpy()
    else:
        _, ctx, (hn, cn) = vae.encode(X, lengths)
        out["context"] = ctx.detach().numpy()
        out["enc_hn"] = hn.detach().numpy()
        out["enc_cn"] = cn.detach().numpy()
    z = torch.cat([output["latent_params"][n].z for n in space_names], dim=1)
    h0, c0 = vae.compute_hidden(z, B)
    out["dec_h0"] = h0.detach().numpy()
    out["dec_c0"] = c0.detach().numpy()

    # losses: run.py:128-163, mi weight 0.01 (run.py:239)
    sys.path.insert(0, "/root/reference")
    L = {}
    L.update(ref_losses.reconstruction_loss(X, output["decoder_logits"], lengths))
    L.update(ref_losses.compute_kl_divergence_losses(vae, output["latent_params"], kl_weights))
    L.update(ref_losses.compute_discriminator_losses(vae, output["dsc_logits"], Y))
    L.update(ref_losses.compute_adversarial_losses(vae, output["adv_logits"], Y))
    L.update(ref_losses.compute_mi_losses(vae, output["latent_params"], beta=0.01))
    total = (L["reconstruction_loss"] + L["total_weighted_kl"] + L["total_dsc_loss"]
             + L["total_adv_loss"] + L["total_mi"])
    out["loss.reconstruction"] = L["reconstruction_loss"].item()
    out["loss.total_weighted_kl"] = L["total_weighted_kl"].item()
    out["loss.total_kl"] = L["total_kl"]
    out["loss.total_dsc"] = L["total_dsc_loss"].item()
    out["loss.total"] = total.item()
    for n, v in L["idv_kls"].items():
        out[f"kl.{n}"] = v
    for n, v in L["idv_dsc_losses"].items():
        out[f"dsc_loss.{n}"] = v
    for n, v in L["idv_dsc_accs"].items():
        out[f"dsc_acc.{n}"] = v

    if mode == "train":
        total.backward(retain_graph=True)
        for k, p in vae.named_parameters():
            if p.grad is not None:
                out[f"grad.{k}"] = p.grad.detach().numpy().copy()
        norm = torch.nn.utils.clip_grad_norm_(vae.trainable_parameters(), 5.0)
        out["grad_norm"] = norm.item()
        opt.step()
        for k, v in vae.state_dict().items():
            out[f"sd_after.{k}"] = v.detach().numpy().copy()

    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB, total loss {total.item():.6f}")


def run_adv_mi_case(name, params, V, label_dims, B, T, kl_weights, seed=10, eps_seed=4321):
    """Full reference train step with the adversarial + MI objectives enabled, in run.py:217-276's order:
    forward, compute_all_losses, total.backward(retain_graph), clip 5.0, adversary optimizer steps, VAE Adam step,
    MI-estimator (CLUB) learning steps.  Records every intermediate the B200 path must reproduce."""
    ref_utils.set_seed(seed)
    vae = ref_model.build_vae(params, V, None, label_dims, torch.device("cpu"), SOS, EOS)
    gen = torch.Generator().manual_seed(seed + 7)
    X, lengths, Y = make_batch(gen, B, T, V, label_dims)
    out = {"B": B, "T": T, "V": V, "sos": SOS, "eos": EOS, "inputs": X.numpy(), "lengths": lengths.numpy(),
           "lr": params["learn_rate"]}
    for k, v in Y.items():
        out[f"Y.{k}"] = v.numpy()
    for k, v in vae.state_dict().items():
        out[f"sd.{k}"] = v.detach().numpy().copy()
    for n, est in vae.mi_estimators.items():            # not part of the state_dict (plain dict, model.py:337)
        for k, v in est.state_dict().items():
            out[f"mi0.{n}.{k}"] = v.detach().numpy().copy()
    space_names = list(vae.context2params.keys())
    out["space_names"] = np.array(space_names)
    out["space_dims"] = np.array([vae.context2params[n].out_features // 2 for n in space_names])
    out["label_names"] = np.array(list(label_dims.keys()))
    out["label_dims"] = np.array(list(label_dims.values()))
    out["adv_names"] = np.array(list(vae.adversaries.keys()))
    out["mi_names"] = np.array(list(vae.mi_estimators.keys()))
    for k in space_names:
        out[f"klw.{k}"] = np.float64(kl_weights.get(k, kl_weights["default"]))
    vae.train()
    torch.manual_seed(eps_seed)
    for n in space_names:
        zs = vae.context2params[n].out_features // 2
        torch.randn(B, zs)
        out[f"eps.{n}"] = torch.randn(B, zs).numpy()
    torch.manual_seed(eps_seed)
    random.seed(seed)
    opt = torch.optim.Adam(vae.trainable_parameters(), lr=params["learn_rate"])
    output = vae(X, lengths, teacher_forcing_prob=1.0)
    for n in space_names:
        out[f"z.{n}"] = output["latent_params"][n].z.detach().numpy()
    for n, l in output["adv_logits"].items():
        out[f"adv_logits.{n}"] = l.detach().numpy()
    L = {}
    L.update(ref_losses.reconstruction_loss(X, output["decoder_logits"], lengths))
    L.update(ref_losses.compute_kl_divergence_losses(vae, output["latent_params"], kl_weights))
    L.update(ref_losses.compute_discriminator_losses(vae, output["dsc_logits"], Y))
    L.update(ref_losses.compute_adversarial_losses(vae, output["adv_logits"], Y))
    L.update(ref_losses.compute_mi_losses(vae, output["latent_params"], beta=0.01))
    total = (L["reconstruction_loss"] + L["total_weighted_kl"] + L["total_dsc_loss"] + L["total_adv_loss"] + L["total_mi"])
    out["loss.total"] = total.item()
    out["loss.total_adv"] = L["total_adv_loss"].item()
    out["loss.total_mi"] = L["total_mi"].item()
    for n, v in L["idv_adv_losses"].items():
        out[f"adv_loss.{n}"] = v
    for n, v in L["idv_adv_dsc_losses"].items():
        out[f"adv_dsc_loss.{n}"] = v.item()
    for n, v in L["idv_adv_dsc_accs"].items():
        out[f"adv_dsc_acc.{n}"] = v
    for n, v in L["idv_mi_estimates"].items():
        out[f"mi_est.{n}"] = v
    total.backward(retain_graph=True)
    for k, p in vae.named_parameters():
        if p.grad is not None:
            out[f"grad.{k}"] = p.grad.detach().numpy().copy()       # adversaries.*: entropy-term gradients only
    out["grad_norm"] = torch.nn.utils.clip_grad_norm_(vae.trainable_parameters(), 5.0).item()
    for n, dl in L["idv_adv_dsc_losses"].items():                  # AdversarialDiscriminator.optimizer_step, unrolled
        adv = vae.adversaries[n]
        dl.backward(retain_graph=True)
        for k, p in adv.named_parameters():
            out[f"adv_grad_at_step.{n}.{k}"] = p.grad.detach().numpy().copy()   # entropy + discriminator gradients
        adv.optimizer.step()
        adv.optimizer.zero_grad()
    opt.step()
    opt.zero_grad()
    for n, est in vae.mi_estimators.items():                       # run.py:264-276
        n1, n2 = n.split('-')
        z1, z2 = output["latent_params"][n1].z.detach(), output["latent_params"][n2].z.detach()
        est.train()
        ll = est.learning_loss(z1, z2)
        out[f"mi_learning_loss.{n}"] = ll.item()
        est.optimizer.zero_grad()
        ll.backward()
        for k, p in est.named_parameters():
            out[f"mi_grad.{n}.{k}"] = p.grad.detach().numpy().copy()
        torch.nn.utils.clip_grad_norm_(est.parameters(), 1.0)
        est.optimizer.step()
        for k, v in est.state_dict().items():
            out[f"mi_after.{n}.{k}"] = v.detach().numpy().copy()
    for k, v in vae.state_dict().items():
        out[f"sd_after.{k}"] = v.detach().numpy().copy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB, total loss {total.item():.6f}, adv {out['loss.total_adv']:.6f}, "
          f"mi {out['loss.total_mi']:.6f}")


def cyclic_table():
    rows = []
    for total in (100, 31260, 7):
        for step in list(range(0, 30)) + [total // 4, total // 4 + 1, total - 1, total // 2 + 3]:
            rows.append((step, total, float(ref_losses.get_cyclic_kl_weight(step, total))))
    path = os.path.join(HERE, "cyclic_kl.npz")
    np.savez_compressed(path, table=np.array(rows, dtype=np.float64))
    print(f"wrote {path}")


def sampler_case():
    """Batches of the reference's RatioSampler (vae/data_utils.py:13-87) on a seeded two-source dataset, for three
    (batch size, ratios) settings -> ratio_sampler.npz (index lists are ragged: stored flat with offsets)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_data_utils", os.path.join(os.environ.get("DVAE_REFERENCE_ROOT", "/root/reference"),
                                                                                 "vae", "data_utils.py"))
    du = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(du)
    rnd = random.Random(0)
    sources = [rnd.choice(["sfu", "amazon", "amazon", "amazon", "amazon"]) for _ in range(203)]
    data = [{"source_dataset": s_} for s_ in sources]
    out = {"sources": np.array(sources)}
    settings = [(16, None), (12, {"sfu": 0.5, "amazon": 0.5}), (10, {"amazon": 0.7, "sfu": 0.3})]
    for i, (bs, ratios) in enumerate(settings):
        torch.manual_seed(3)
        sampler = du.RatioSampler(data, "source_dataset", ratios, bs)
        batches = [b.numpy() for b in sampler]
        out[f"case{i}.batch_size"] = np.int64(bs)
        out[f"case{i}.ratio_keys"] = np.array(list(ratios) if ratios else [], dtype=str)
        out[f"case{i}.ratio_vals"] = np.array(list(ratios.values()) if ratios else [], dtype=np.float64)
        out[f"case{i}.flat"] = np.concatenate(batches)
        out[f"case{i}.sizes"] = np.array([len(b) for b in batches], dtype=np.int64)
        out[f"case{i}.len"] = np.int64(len(sampler))
    path = os.path.join(HERE, "ratio_sampler.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}")


def formats_case():
    """On-disk formats: the metadata logs `run.log_params` writes (run.py:166-194) for a seeded dict, as file-name ->
    content strings, and the key structure of the reference's checkpoint (run.py:624-630) -> formats.npz."""
    import importlib
    import tempfile
    sys.path.insert(0, os.environ.get("DVAE_REFERENCE_ROOT", "/root/reference"))
    ref_run = importlib.import_module("run")
    rng = np.random.default_rng(5)
    params = {"polarity": {"z": rng.standard_normal((7, 1)).tolist(), "mu": rng.standard_normal((7, 1)).tolist()},
              "content": {"z": (rng.standard_normal((7, 3)) * 10).tolist(), "logvar": rng.standard_normal((7, 3)).tolist()}}
    ids = [f"ex{i}" for i in range(7)]
    out = {"ids": np.array(ids)}
    for ln, d_ in params.items():
        for pn, v in d_.items():
            out[f"in.{ln}.{pn}"] = np.array(v)
    with tempfile.TemporaryDirectory() as td:
        ref_run.log_params(params, ids, td, "train", 4)
        files = {}
        for root, _, fs in os.walk(td):
            for f in fs:
                files[os.path.relpath(os.path.join(root, f), td)] = open(os.path.join(root, f), newline="").read()
    out["file_names"] = np.array(sorted(files))
    out["file_contents"] = np.array([files[k] for k in sorted(files)])
    # checkpoint structure: one Adam step on a tiny reference model
    torch.manual_seed(10)
    p = make_params()
    vae = ref_model.build_vae(p, 23, None, {"polarity": 1}, torch.device("cpu"), SOS, EOS)
    opt = torch.optim.Adam(vae.trainable_parameters(), lr=p["learn_rate"])
    sum(q.sum() for q in vae.trainable_parameters()).backward()
    opt.step()
    sd = opt.state_dict()
    out["opt.param_group_keys"] = np.array(sorted(sd["param_groups"][0].keys()))
    out["opt.state_keys"] = np.array(sorted(sd["state"][0].keys()))
    out["opt.n_params"] = np.int64(len(sd["param_groups"][0]["params"]))
    out["opt.trainable_names"] = np.array([n for n, q in vae.named_parameters()
                                           if q.requires_grad and not n.startswith("adversaries")])
    path = os.path.join(HERE, "formats.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}")


if __name__ == "__main__":
    torch.set_num_threads(1)
    if "--sampler" in sys.argv:
        sampler_case()
        sys.exit(0)
    if "--formats" in sys.argv:
        formats_case()
        sys.exit(0)
    if "--bow" in sys.argv:             # only the BOWEncoder case (vae/model.py:13-49)
        run_case("tiny_bow", make_params(bow_encoder=True, embedding_dim=10, hidden_dim=8,
                                         latent_dims={"total": 5, "polarity": 1}),
                 V=27, label_dims={"polarity": 1}, B=5, T=7, kl_weights={"default": 0.1, "polarity": 0.005})
        sys.exit(0)
    if "--adv-mi" in sys.argv:          # only the adversarial + MI case (leaves the other files untouched)
        run_adv_mi_case("tiny_adv_mi", make_params(bidirectional_encoder=True, embedding_dim=10, hidden_dim=8,
                                                    latent_dims={"total": 9, "polarity": 1, "uncertainty": 2},
                                                    learn_rate=3e-4, adversarial_loss=True, mi_loss=True,
                                                    lambdas={"default": 0.2, "polarity": 0.005, "uncertainty": 0.005}),
                        V=31, label_dims={"uncertainty": 3, "polarity": 1}, B=6, T=7,
                        kl_weights={"default": 0.2, "polarity": 0.005, "uncertainty": 0.005})
        sys.exit(0)
    # cfg-1-like: uni-directional encoder, polarity + content (config_example.json shape, shrunk)
    run_case("tiny_uni", make_params(), V=37, label_dims={"polarity": 1}, B=5, T=8,
             kl_weights={"default": 0.01, "polarity": 0.005})
    # cfg-2-like: bi-directional encoder, uncertainty + polarity + content, ragged lengths;
    # the clip (5.0) is made to bite by a large lr-independent loss scale via short rows
    run_case("tiny_bi", make_params(bidirectional_encoder=True, embedding_dim=10, hidden_dim=8,
                                    latent_dims={"total": 7, "polarity": 1, "uncertainty": 2},
                                    learn_rate=3e-4,
                                    lambdas={"default": 0.37, "polarity": 0.005, "uncertainty": 0.005}),
             V=29, label_dims={"uncertainty": 1, "polarity": 1}, B=6, T=9,
             kl_weights={"default": 0.37, "polarity": 0.005, "uncertainty": 0.005})
    # eval mode (single eps draw, model.py:393-395), multi-class label head, 1-layer config
    # (decoder silently gets 2 layers: model.py:123-124)
    run_case("tiny_eval_mc", make_params(num_rnn_layers=1, embedding_dim=6, hidden_dim=8,
                                         latent_dims={"total": 5, "modality": 2}),
             V=23, label_dims={"modality": 3}, B=4, T=6,
             kl_weights={"default": 1.0}, mode="eval")
    cyclic_table()
