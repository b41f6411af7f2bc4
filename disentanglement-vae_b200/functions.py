"""Autograd glue between the drop-in module surface (`model.py`, `losses.py`) and the kernel plan
(`plan.py`).  Each `torch.autograd.Function` here wraps C-ABI calls for forward and backward;
nothing in this file computes on the path with torch ops.
"""
import weakref
from collections import OrderedDict

import torch
from torch.autograd import Function

from . import _lib
from ._lib import ptr, check, int_array
from .model import Params, LatentParams, DscLogits, FusedLogits


class _PlanToken:
    """Releases the plan's buffers when the autograd graph that uses them is dropped."""

    def __init__(self, plan):
        self._fin = weakref.finalize(self, _PlanToken._release, plan)

    @staticmethod
    def _release(plan):
        plan.busy = False


def _c(t):
    return None if t is None else t.contiguous()


def _param_list(model):
    named = dict(model.named_parameters())
    names = list(model._layout.keys())
    return names, [named[n] for n in names]


# --------------------------------------------------------------------------------------------
# encoder -> fused heads -> teacher-forced decoder
# --------------------------------------------------------------------------------------------
class _EncDecFn(Function):
    @staticmethod
    def forward(ctx, model, plan, inputs, lengths, eps, dec_tokens, first_token, train, *params):
        """`first_token`: int = <SOS> id for the teacher-forced whole-sequence decoder; a tuple of per-step
        coins = sampled decoding (sampling.py), where dec_tokens is the [B,T] prediction buffer."""
        P = model._P
        plan.encode(P, inputs, lengths, train)
        plan.heads(P, plan.ctx, eps, None, None)
        if isinstance(first_token, tuple):
            plan.decode_sampled(P, dec_tokens, first_token, train)
        else:
            plan.decode_forced(P, dec_tokens, first_token, train)
        ctx.model, ctx.plan, ctx.token = model, plan, _PlanToken(plan)
        ctx.inputs, ctx.lengths, ctx.eps = inputs, lengths, eps
        S = plan.d.S
        return (plan.d_hs[-1].detach(), plan.z.clone(), plan.mu.clone(), plan.logvar.clone(),
                plan.dsc_logits.clone(), plan.scalars[3:3 + S].clone(), plan.ctx.clone())

    @staticmethod
    def backward(ctx, g_htop, g_z, g_mu, g_logvar, g_dsc, g_kl, g_ctx):
        model, plan = ctx.model, ctx.plan
        P = model._P
        flat_grad = torch.empty(model._flat_numel, device=plan.device, dtype=torch.float32)
        G = model.grad_views(flat_grad)
        enc_emb = "encoder.embedding.weight" in model._layout
        dec_emb = "decoder.embedding.weight" in model._layout
        if enc_emb:
            G["encoder.embedding.weight"].zero_()
        if dec_emb:
            G["decoder.embedding.weight"].zero_()
        if g_htop is None:
            g_htop = torch.zeros_like(plan.d_hs[-1])
        lib = _lib.load()
        check(lib.dvae_defer_joins(1), "dvae_defer_joins")      # weight-gradient GEMMs overlap the next layer's recurrence
        try:
            g_hid = plan.decode_bwd(P, G, _c(g_htop), emb_grad=dec_emb)
            kl_w = _c(g_kl)
            g_c = plan.heads_bwd(P, G, plan.ctx, ctx.eps, None, kl_w, g_hid, _c(g_z), _c(g_mu), _c(g_logvar), _c(g_dsc))
            if g_ctx is not None:
                g_c.add_(g_ctx)
            plan.encode_bwd(P, G, ctx.inputs, ctx.lengths, g_c, emb_grad=enc_emb)
        finally:
            check(lib.dvae_join_side_streams(_lib.stream_ptr()), "dvae_join_side_streams")
        plan.busy = False
        names = list(model._layout.keys())
        skip = ("decoder.linear.weight", "decoder.linear.bias")
        return (None,) * 8 + tuple(None if n in skip else G[n] for n in names)


# --------------------------------------------------------------------------------------------
# fused vocabulary projection + cross entropy
# --------------------------------------------------------------------------------------------
class _VocabCEFn(Function):
    @staticmethod
    def forward(ctx, plan, h_top, targets, lengths, weight, bias):
        P = {"decoder.linear.weight": weight, "decoder.linear.bias": bias}
        plan.vocab_ce(P, h_top, targets, lengths)
        ctx.plan, ctx.h_top, ctx.targets, ctx.lengths = plan, h_top, targets, lengths
        ctx.save_for_backward(weight, bias)
        return plan.recon.clone().reshape(())

    @staticmethod
    def backward(ctx, g_loss):
        plan = ctx.plan
        weight, bias = ctx.saved_tensors
        P = {"decoder.linear.weight": weight, "decoder.linear.bias": bias}
        G = {"decoder.linear.weight": torch.empty_like(weight), "decoder.linear.bias": torch.empty_like(bias)}
        gs = g_loss.reshape(1).to(torch.float32).contiguous()
        g_top = plan.vocab_ce_bwd(P, G, ctx.h_top, ctx.targets, ctx.lengths, gs)
        return None, g_top.detach(), None, None, G["decoder.linear.weight"], G["decoder.linear.bias"]


class _LogitsFn(Function):
    """Dense logits [T1*B, V] = h_top W^T + b (debug / evaluation only)."""

    @staticmethod
    def forward(ctx, h_top, weight, bias):
        lib = _lib.load()
        N, H, V = h_top.numel() // h_top.size(-1), h_top.size(-1), weight.size(0)
        out = torch.empty(N, V, device=h_top.device, dtype=torch.float32)
        check(lib.dvae_linear(ptr(h_top), H, 0, ptr(weight), H, 0, ptr(out), V, N, V, H, ptr(bias), None, 0.0, 0,
                              _lib.stream_ptr()), "dvae_linear")
        ctx.save_for_backward(h_top, weight)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        h_top, weight = ctx.saved_tensors
        g = g.contiguous()
        N, V, H = g.size(0), g.size(1), weight.size(1)
        st = _lib.stream_ptr()
        d_h = torch.empty_like(h_top)
        d_w = torch.empty_like(weight)
        d_b = torch.empty(V, device=g.device, dtype=torch.float32)
        check(lib.dvae_linear(ptr(g), V, 0, ptr(weight), H, 1, ptr(d_h), H, N, H, V, None, None, 0.0, 0, st), "dvae_linear")
        check(lib.dvae_linear(ptr(g), V, 1, ptr(h_top), H, 1, ptr(d_w), H, V, H, N, None, None, 0.0, 0, st), "dvae_linear")
        check(lib.dvae_colsum(ptr(g), V, N, V, ptr(d_b), 0.0, st), "dvae_colsum")
        return d_h, d_w, d_b


class _SmallLinearFn(Function):
    """y = act(x W^T + b) for the small heads on the latent spaces (adversaries, CLUB estimators).
    act: 0 none, 1 tanh, 2 ReLU.  x may be a column slice of a wider matrix (row stride = x.stride(0))."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        lib = _lib.load()
        if x.stride(-1) != 1:
            x = x.contiguous()
        Bn, K, N = x.size(0), x.size(1), weight.size(0)
        st = _lib.stream_ptr()
        out = torch.empty(Bn, N, device=x.device, dtype=torch.float32)
        check(lib.dvae_linear(ptr(x), x.stride(0), 0, ptr(weight), K, 0, ptr(out), N, Bn, N, K, ptr(bias), None, 0.0,
                              1 if act == 1 else 0, st), "dvae_linear")
        if act == 2:
            check(lib.dvae_relu(ptr(out), out.numel(), st), "dvae_relu")
        ctx.save_for_backward(x, weight, out)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        x, weight, out = ctx.saved_tensors
        Bn, K, N = x.size(0), x.size(1), weight.size(0)
        st = _lib.stream_ptr()
        g = g.contiguous()
        if ctx.act:
            d = torch.empty_like(g)
            check(lib.dvae_act_bwd(ptr(out), ptr(g), ptr(d), g.numel(), ctx.act, st), "dvae_act_bwd")
            g = d
        d_x = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            d_x = torch.empty(Bn, K, device=g.device, dtype=torch.float32)
            check(lib.dvae_linear(ptr(g), N, 0, ptr(weight), K, 1, ptr(d_x), K, Bn, K, N, None, None, 0.0, 0, st), "dvae_linear")
        if ctx.needs_input_grad[1]:
            d_w = torch.empty_like(weight)
            check(lib.dvae_linear(ptr(g), N, 1, ptr(x), x.stride(0), 1, ptr(d_w), K, N, K, Bn, None, None, 0.0, 0, st), "dvae_linear")
        if ctx.needs_input_grad[2]:
            d_b = torch.empty(N, device=g.device, dtype=torch.float32)
            check(lib.dvae_colsum(ptr(g), N, Bn, N, ptr(d_b), 0.0, st), "dvae_colsum")
        return d_x, d_w, d_b, None


class _EntropyLossFn(Function):
    """mean_b sum_c p log p of clamped sigmoid / softmax probabilities (vae/model.py:247-258)."""

    @staticmethod
    def forward(ctx, logits):
        lib = _lib.load()
        logits = logits.contiguous()
        loss = torch.empty(1, device=logits.device, dtype=torch.float32)
        check(lib.dvae_entropy_loss(ptr(logits), logits.size(0), logits.size(1), ptr(loss), None, None, _lib.stream_ptr()),
              "dvae_entropy_loss")
        ctx.save_for_backward(logits)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (logits,) = ctx.saved_tensors
        d = torch.empty_like(logits)
        gs = g.reshape(1).to(torch.float32).contiguous()
        check(lib.dvae_entropy_loss(ptr(logits), logits.size(0), logits.size(1), None, ptr(gs), ptr(d), _lib.stream_ptr()),
              "dvae_entropy_loss")
        return d


class _ClubMiFn(Function):
    """CLUB.forward's estimate from (mu, logvar) = q(y|x) and the samples y (vae/losses.py:53-67)."""

    @staticmethod
    def forward(ctx, mu, logvar, y):
        lib = _lib.load()
        mu, logvar, y = mu.contiguous(), logvar.contiguous(), y.contiguous()
        out = torch.empty(1, device=mu.device, dtype=torch.float32)
        check(lib.dvae_club_mi(ptr(mu), ptr(logvar), ptr(y), mu.size(0), mu.size(1), ptr(out), None, None, None, None, None,
                               _lib.stream_ptr()), "dvae_club_mi")
        ctx.save_for_backward(mu, logvar, y)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        mu, logvar, y = ctx.saved_tensors
        d_mu, d_lv, d_y = torch.empty_like(mu), torch.empty_like(mu), torch.empty_like(mu)
        ws = torch.empty(3 * mu.size(1), device=mu.device, dtype=torch.float32)
        gs = g.reshape(1).to(torch.float32).contiguous()
        check(lib.dvae_club_mi(ptr(mu), ptr(logvar), ptr(y), mu.size(0), mu.size(1), None, ptr(gs), ptr(d_mu), ptr(d_lv),
                               ptr(d_y), ptr(ws), _lib.stream_ptr()), "dvae_club_mi")
        return d_mu, d_lv, d_y


class _ClubNllFn(Function):
    """CLUB.learning_loss = -loglikeli (vae/losses.py:69-74)."""

    @staticmethod
    def forward(ctx, mu, logvar, y):
        lib = _lib.load()
        mu, logvar, y = mu.contiguous(), logvar.contiguous(), y.contiguous()
        out = torch.empty(1, device=mu.device, dtype=torch.float32)
        check(lib.dvae_club_nll(ptr(mu), ptr(logvar), ptr(y), mu.size(0), mu.size(1), ptr(out), None, None, None,
                                _lib.stream_ptr()), "dvae_club_nll")
        ctx.save_for_backward(mu, logvar, y)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        mu, logvar, y = ctx.saved_tensors
        d_mu, d_lv = torch.empty_like(mu), torch.empty_like(mu)
        gs = g.reshape(1).to(torch.float32).contiguous()
        check(lib.dvae_club_nll(ptr(mu), ptr(logvar), ptr(y), mu.size(0), mu.size(1), None, ptr(gs), ptr(d_mu), ptr(d_lv),
                                _lib.stream_ptr()), "dvae_club_nll")
        return d_mu, d_lv, None


class _DscLossFn(Function):
    """Per-discriminator loss and accuracy from packed logits (vae/losses.py:180-196)."""

    @staticmethod
    def forward(ctx, packed_logits, labels, space_dims, dsc_out):
        lib = _lib.load()
        S, B = len(space_dims), packed_logits.size(0)
        out = torch.zeros(2 * S, device=packed_logits.device, dtype=torch.float32)
        check(lib.dvae_dsc_loss(ptr(packed_logits), ptr(labels), B, S, int_array(space_dims), int_array(dsc_out),
                                ptr(out), None, None, _lib.stream_ptr()), "dvae_dsc_loss")
        ctx.save_for_backward(packed_logits, labels)
        ctx.meta = (space_dims, dsc_out)
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _lib.load()
        packed_logits, labels = ctx.saved_tensors
        space_dims, dsc_out = ctx.meta
        S, B = len(space_dims), packed_logits.size(0)
        d_logits = torch.zeros_like(packed_logits)
        check(lib.dvae_dsc_loss(ptr(packed_logits), ptr(labels), B, S, int_array(space_dims), int_array(dsc_out),
                                None, ptr(g_out.contiguous()), ptr(d_logits), _lib.stream_ptr()), "dvae_dsc_loss")
        return d_logits, None, None, None


# --------------------------------------------------------------------------------------------
# helpers behind the module methods
# --------------------------------------------------------------------------------------------
def _prep_tokens(model, inputs, lengths):
    dev = model._flat.device
    if inputs.dim() != 2:
        raise ValueError(f"inputs must be [batch, length], got {tuple(inputs.shape)}")
    inputs = inputs.to(device=dev, dtype=torch.int64)
    if inputs.stride(1) != 1:
        inputs = inputs.contiguous()
    if not torch.is_tensor(lengths):
        lengths = torch.tensor(lengths, dtype=torch.int64)
    lengths = lengths.to(device=dev, dtype=torch.int64).contiguous()
    if lengths.numel() != inputs.size(0):
        raise ValueError("lengths must have one entry per batch row")
    return inputs, lengths


def _set_seed(model, plan):
    d = plan.d
    if model.training and (d.p_enc > 0.0 or d.p_dec > 0.0):
        # host-side draw from torch's (seeded) CPU generator: no device sync
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        plan.seed_dev.fill_(seed)


def _package(model, plan, outs, B, T, inputs):
    h_top, z, mu, logvar, dsc, kl, context = outs
    d = plan.d
    lat = LatentParams()
    off = 0
    for n, zs in zip(d.space_names, d.space_dims):
        lat[n] = Params(z[:, off:off + zs], mu[:, off:off + zs], logvar[:, off:off + zs])
        off += zs
    lat.kl, lat.names = kl, tuple(d.space_names)
    logits = DscLogits()
    off = 0
    for n, o in zip(d.space_names, d.dsc_out):
        if o > 0:
            logits[n] = dsc[:, off:off + o]
            off += o
    logits.packed, logits.dims = dsc, (tuple(d.space_dims), tuple(d.dsc_out))
    return lat, logits, context


def run_forward(model, inputs, lengths, coins, eps=None):
    model._require_cuda()
    inputs, lengths = _prep_tokens(model, inputs, lengths)
    B, T = inputs.shape
    if T < 2:
        raise ValueError("sequences must have at least 2 positions (<SOS> plus one target)")
    if not all(coins):
        from .sampling import run_forward_sampled
        return run_forward_sampled(model, inputs, lengths, coins, eps)
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    plan = model.get_plan(B, T, need_grad)
    _set_seed(model, plan)
    if eps is None:
        eps = model._draw_eps(B, inputs.device)
    eps = eps.to(device=inputs.device, dtype=torch.float32).contiguous()
    _, params = _param_list(model)
    outs = _EncDecFn.apply(model, plan, inputs, lengths, eps, inputs, model.sos_token_idx, model.training, *params)
    lat, dsc, context = _package(model, plan, outs, B, T, inputs)
    preds = inputs.clone()
    preds[:, 0] = model.sos_token_idx          # tf: predictions are the forced next inputs (model.py:464-472)
    logits = FusedLogits(model, plan, outs[0], B, T)
    return {"decoder_logits": logits, "latent_params": lat, "dsc_logits": dsc, "adv_logits": adversary_logits(model, lat),
            "token_predictions": preds, "context": context}


def adversary_logits(model, latent_params):
    """vae/model.py:432-436: every adversary reads the sampled z of its latent space."""
    out = {}
    for name, adv in model.adversaries.items():
        out[name] = adv(latent_params[name.split('-')[0]].z)
    return out


def run_encoder(model, inputs, lengths):
    """encode() for inspection callers (vae/model.py:373-382); forward only."""
    model._require_cuda()
    if model._dims.bow:
        raise _lib.DvaeError("encode() is the recurrent encoder's inspection API (vae/model.py:373-382); the BOW encoder has "
                             "no hidden states -- use forward()")
    inputs, lengths = _prep_tokens(model, inputs, lengths)
    B, T = inputs.shape
    plan = model.get_plan(B, T, False)
    _set_seed(model, plan)
    with torch.no_grad():
        plan.encode(model._P, inputs, lengths, model.training)
        d = plan.d
        tmax = int(lengths.max())
        encoded = plan.e_hs[-1][:tmax].transpose(0, 1).clone()
        context = plan.ctx.clone()
        hn = torch.stack([plan.ctx[:, k * d.H:(k + 1) * d.H] for k in range(d.Le * d.D)]).clone()
        cn = torch.stack([plan.ctx_c[:, k * d.H:(k + 1) * d.H] for k in range(d.Le * d.D)]).clone()
    return encoded, context, (hn, cn)


def run_heads(model, context, eps=None):
    """compute_latent_params() on a caller-supplied context (vae/model.py:384-398); forward only."""
    model._require_cuda()
    context = context.to(device=model._flat.device, dtype=torch.float32).contiguous()
    B = context.size(0)
    plan = model.get_plan(B, 2, False)
    if eps is None:
        eps = model._draw_eps(B, context.device)
    eps = eps.to(device=context.device, dtype=torch.float32).contiguous()
    with torch.no_grad():
        plan.heads(model._P, context, eps, None, None)
        S = plan.d.S
        outs = (None, plan.z.clone(), plan.mu.clone(), plan.logvar.clone(), plan.dsc_logits.clone(),
                plan.scalars[3:3 + S].clone(), context)
    return _package(model, plan, outs, B, 2, None)


def run_z2hidden(model, z):
    """compute_hidden() (vae/model.py:400-411): tanh(z2hidden(z)) split into (state, cell) [Ld,B,H]."""
    model._require_cuda()
    lib = _lib.load()
    z = z.to(device=model._flat.device, dtype=torch.float32).contiguous()
    B, Z = z.shape
    P = model._P
    H, Ld = model.decoder.hidden_size, model.decoder.num_layers
    hid = torch.empty(B, 2 * H * Ld, device=z.device, dtype=torch.float32)
    check(lib.dvae_linear(ptr(z), Z, 0, ptr(P["z2hidden.weight"]), Z, 0, ptr(hid), 2 * H * Ld, B, 2 * H * Ld, Z,
                          ptr(P["z2hidden.bias"]), None, 0.0, 1, _lib.stream_ptr()), "dvae_linear")
    state = hid[:, :H * Ld].reshape(B, Ld, H).transpose(0, 1)
    cell = hid[:, H * Ld:].reshape(B, Ld, H).transpose(0, 1)
    return state, cell


def run_sample(model, z, max_length):
    from .sampling import run_sample as _rs
    return _rs(model, z, max_length)


def reconstruction_loss_fused(logits, targets, lengths):
    model, plan = logits.model, logits.plan
    targets, lengths = _prep_tokens(model, targets, lengths)
    named = dict(model.named_parameters())
    loss = _VocabCEFn.apply(plan, logits.h_top, targets, lengths, named["decoder.linear.weight"],
                            named["decoder.linear.bias"])
    B, T = targets.shape
    am = torch.empty(B, T, device=targets.device, dtype=torch.int64)
    am[:, 0] = model.sos_token_idx
    am[:, 1:] = plan.argmax[:plan.N].view(T - 1, B).t()
    logits.argmax_tokens = am
    return loss


def vocab_argmax(logits):
    if logits.argmax_tokens is None:
        model, plan = logits.model, logits.plan
        B, T, _ = logits.shape
        dummy_t = torch.zeros(B, T, device=plan.device, dtype=torch.int64)
        dummy_l = torch.full((B,), T, device=plan.device, dtype=torch.int64)
        with torch.no_grad():
            reconstruction_loss_fused(logits, dummy_t, dummy_l)
    return logits.argmax_tokens


def materialize_logits(logits):
    model = logits.model
    named = dict(model.named_parameters())
    B, T, V = logits.shape
    flat = _LogitsFn.apply(logits.h_top, named["decoder.linear.weight"], named["decoder.linear.bias"])
    out = torch.zeros(B, T, V, device=flat.device, dtype=torch.float32)
    out[:, 0, model.sos_token_idx] = 1.0                  # vae/model.py:454
    out[:, 1:, :] = flat.view(T - 1, B, V).transpose(0, 1)
    return out
