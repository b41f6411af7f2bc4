"""Sampled decoding: `forward(..., teacher_forcing_prob < 1)` (vae/model.py:463-472) and `sample()`
(vae/model.py:484-512).

The reference draws the next decoder input with `torch.multinomial(softmax(logits))` from a dense
`[B,V]` logits row.  Here the decoder advances one step per C-ABI call (`dvae_lstm_step`) and the
next token comes from `dvae_vocab_sample_step`: vocabulary projection + Gumbel-max in one kernel
(arg-max of logits + Gumbel(0,1) noise is a draw from softmax(logits)), so neither logits nor
probabilities are written to HBM.  The step calls leave the decoder buffers exactly as a
whole-sequence call would, so the loss and the whole backward pass are shared with the
teacher-forced path (sampled tokens are discrete: no gradient flows through them, as in the
reference).

RNG: the draws are Philox-keyed by (step seed, decode step, row, vocabulary column), not by
torch's generator, so the *distribution* matches the reference but not its random stream.
`token_predictions` stays on the device (the reference builds it on the CPU, vae/model.py:455);
callers' `.to(device)` / `.cpu()` work unchanged.
"""
import torch

from . import _lib
from ._lib import ptr, check
from .model import FusedLogits


def _seed_sampling(plan):
    # host-side draw from torch's (seeded) CPU generator: no device sync
    plan.seed_dev.fill_(int(torch.randint(0, 2 ** 62, (1,)).item()))


def run_forward_sampled(model, inputs, lengths, coins, eps=None):
    """forward() when at least one decoding step samples its input (coins[i-1] False for position i)."""
    from .functions import _EncDecFn, _param_list, _package, adversary_logits
    B, T = inputs.shape
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    plan = model.get_plan(B, T, need_grad)
    _seed_sampling(plan)
    if eps is None:
        eps = model._draw_eps(B, inputs.device)
    eps = eps.to(device=inputs.device, dtype=torch.float32).contiguous()
    preds = inputs.clone()                      # forced positions keep inputs[:, i]; sampled ones are overwritten
    preds[:, 0] = model.sos_token_idx
    _, params = _param_list(model)
    outs = _EncDecFn.apply(model, plan, inputs, lengths, eps, preds, tuple(coins), model.training, *params)
    lat, dsc, context = _package(model, plan, outs, B, T, inputs)
    logits = FusedLogits(model, plan, outs[0], B, T)
    return {"decoder_logits": logits, "latent_params": lat, "dsc_logits": dsc, "adv_logits": adversary_logits(model, lat),
            "token_predictions": preds, "context": context}


def run_sample(model, z, max_length):
    """sample(): decode `max_length` positions from latent z with every input sampled; forward only."""
    model._require_cuda()
    lib = _lib.load()
    z = z.to(device=model._flat.device, dtype=torch.float32).contiguous()
    if z.dim() != 2 or z.size(1) != model.latent_dim:
        raise ValueError(f"z must be [batch, {model.latent_dim}], got {tuple(z.shape)}")
    if max_length < 2:
        raise ValueError("max_length must be at least 2 (<SOS> plus one sampled position)")
    B = z.size(0)
    plan = model.get_plan(B, max_length, False)
    _seed_sampling(plan)
    P, d = model._P, plan.d
    preds = torch.zeros(B, max_length, device=z.device, dtype=torch.int64)
    preds[:, 0] = model.sos_token_idx
    with torch.no_grad():
        # hid = tanh(z2hidden(z)) (vae/model.py:400-411), consumed by the decoder through row strides
        check(lib.dvae_linear(ptr(z), d.Z, 0, ptr(P["z2hidden.weight"]), d.Z, 0, ptr(plan.hid), d.H2L, B, d.H2L,
                              d.Z, ptr(P["z2hidden.bias"]), None, 0.0, 1, _lib.stream_ptr()), "dvae_linear")
        h_top = plan.decode_sampled(P, preds, [False] * (max_length - 1), model.training)
    logits = FusedLogits(model, plan, h_top.detach().clone(), B, max_length)
    return {"decoder_logits": logits, "token_predictions": preds}
