"""Drop-in replacement for `vae/losses.py` (loss surface of jvasilakes/disentanglement-vae).

Same function names, arguments and returned dict keys (vae/losses.py:137-242).  When handed the
outputs of this package's `VariationalSeq2Seq.forward` the heavy terms come straight from the
fused kernels: the reconstruction loss from the vocab-CE kernel (logits never materialised), the
per-space KL from the fused-heads kernel, the discriminator losses from `dvae_dsc_loss`.
"""
import math

import torch

from . import _lib
from .model import FusedLogits, LatentParams, DscLogits


def reconstruction_loss(targets, logits, target_lengths):
    """vae/losses.py:137-140 (texar sequence_sparse_softmax_cross_entropy defaults: sum over
    time, mean over batch, positions t < length)."""
    if not isinstance(logits, FusedLogits):
        raise _lib.DvaeError("reconstruction_loss expects the `decoder_logits` returned by the B200 model "
                             "(a FusedLogits); dense logits are only available via .materialize()")
    from .functions import reconstruction_loss_fused
    return {"reconstruction_loss": reconstruction_loss_fused(logits, targets, target_lengths)}


def get_cyclic_kl_weight(step, total_steps, cycles=4, rate=0.5):
    """vae/losses.py:143-150: linear ramp over the first `rate` of each of `cycles` periods."""
    period = total_steps / cycles
    tau = (step % math.ceil(period)) / period
    return tau / rate if tau <= rate else 1


def kl_divergence(mu, logvar):
    """vae/losses.py:153-156 on arbitrary tensors (not on the fused path; kept for API parity)."""
    return (0.5 * (torch.exp(logvar) + mu * mu - 1 - logvar)).mean(0).sum()


def compute_kl_divergence_losses(model, latent_params, kl_weights_dict):
    """vae/losses.py:159-177.  One host read for all per-space values instead of one per space."""
    names = list(latent_params.keys())
    if isinstance(latent_params, LatentParams) and latent_params.kl is not None:
        kl_vec = latent_params.kl
    else:
        kl_vec = torch.stack([kl_divergence(p.mu, p.logvar) for p in latent_params.values()])
    weights = [kl_weights_dict[n] if n in kl_weights_dict else kl_weights_dict["default"] for n in names]
    w = torch.tensor(weights, dtype=torch.float32).to(kl_vec.device, non_blocking=True)
    total_weighted_kl = (kl_vec * w).sum()
    vals = kl_vec.detach().tolist()
    idv_kls = dict(zip(names, vals))
    return {"total_weighted_kl": total_weighted_kl, "total_kl": float(sum(vals)), "idv_kls": idv_kls}


def _pack_labels(names, Ybatch, device):
    rows = [Ybatch[n].to(device=device, dtype=torch.float32).reshape(-1) for n in names]
    return torch.stack(rows).contiguous()


def compute_discriminator_losses(model, discriminator_logits, Ybatch):
    """vae/losses.py:180-196."""
    names = list(discriminator_logits.keys())
    device = model.device
    if not names:
        return {"total_dsc_loss": torch.tensor(0.0, device=device), "idv_dsc_losses": {}, "idv_dsc_accs": {}}
    from .functions import _DscLossFn
    if isinstance(discriminator_logits, DscLogits) and discriminator_logits.packed is not None:
        packed, (space_dims, dsc_out) = discriminator_logits.packed, discriminator_logits.dims
    else:
        packed = torch.cat([discriminator_logits[n] for n in names], dim=1).contiguous()
        space_dims = tuple(model.discriminators[n].latent_dim for n in names)
        dsc_out = tuple(model.discriminators[n].output_dim for n in names)
    labels = _pack_labels(names, Ybatch, packed.device)
    out = _DscLossFn.apply(packed, labels, space_dims, dsc_out)
    S = len(space_dims)
    idx = [i for i, o in enumerate(dsc_out) if o > 0]
    sel = torch.tensor(idx, device=out.device)
    total = out[:S].index_select(0, sel).sum()
    vals = out.detach().tolist()
    return {"total_dsc_loss": total,
            "idv_dsc_losses": {n: vals[i] for n, i in zip(names, idx)},
            "idv_dsc_accs": {n: vals[S + i] for n, i in zip(names, idx)}}


def _single_dsc_loss(dsc, logits, targets):
    from .functions import _DscLossFn
    labels = targets.to(device=logits.device, dtype=torch.float32).reshape(1, -1).contiguous()
    out = _DscLossFn.apply(logits.contiguous(), labels, (dsc.latent_dim,), (dsc.output_dim,))
    return out[0], out[1]


def compute_adversarial_losses(model, adversary_logits, Ybatch):
    """vae/losses.py:199-223 with no adversaries configured (the accelerated path's scope)."""
    if len(adversary_logits) > 0:
        raise NotImplementedError("adversarial objective: SURVEY.md 8f n2")
    return {"total_adv_loss": torch.tensor(0.0, device=model.device), "idv_adv_losses": {},
            "idv_adv_dsc_losses": {}, "idv_adv_dsc_accs": {}}


def compute_mi_losses(model, latent_params, beta=1.0):
    """vae/losses.py:226-242 with no MI estimators configured."""
    if len(model.mi_estimators) > 0:
        raise NotImplementedError("MI objective: SURVEY.md 8f n2")
    return {"total_mi": torch.tensor(0.0, device=model.device), "idv_mi_estimates": {}}


def compute_all_losses(model, model_outputs, Xbatch, Ybatch, lengths, kl_weights_dict, mi_loss_weight=0.01):
    """run.py:128-163."""
    L = dict()
    for part in (reconstruction_loss(Xbatch, model_outputs["decoder_logits"], lengths),
                 compute_kl_divergence_losses(model, model_outputs["latent_params"], kl_weights_dict),
                 compute_discriminator_losses(model, model_outputs["dsc_logits"], Ybatch),
                 compute_adversarial_losses(model, model_outputs["adv_logits"], Ybatch),
                 compute_mi_losses(model, model_outputs["latent_params"], beta=mi_loss_weight)):
        for k, v in part.items():
            L.setdefault(k, v)
    total = (L["reconstruction_loss"] + L["total_weighted_kl"] + L["total_dsc_loss"] + L["total_adv_loss"]
             + L["total_mi"])
    return total, L
