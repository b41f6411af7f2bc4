"""Drop-in replacement for `vae/model.py` of jvasilakes/disentanglement-vae, backed by the
hand-written sm_100a kernels behind `include/dvae_b200.h`.

Same factory (`build_vae`, vae/model.py:515-559), same module / attribute surface, same
`state_dict()` keys and shapes, same output-dict schema (vae/model.py:477-481) -- see
INTEGRATION.md.  The torch `nn.LSTM` / `nn.Linear` / `nn.Embedding` objects below are PARAMETER
CONTAINERS ONLY (identical names, shapes and seeded initialisation as the reference); their
`forward` is never called.  All arithmetic on the path runs in the CUDA extension; there is no
CPU or eager-PyTorch fallback.
"""
import random
import weakref
from collections import namedtuple, OrderedDict

import torch
import torch.nn as nn

from . import _lib
from .plan import StepPlan, Dims

Params = namedtuple("Params", ["z", "mu", "logvar"])


class _DeviceMixin:
    @property
    def device(self):
        return self._device

    def set_device(self, value):
        assert isinstance(value, torch.device)
        self._device = value
        self.to(value)


class BOWEncoder(_DeviceMixin, nn.Module):
    """Bag-of-words encoder (vae/model.py:13-49): context = max over positions of dropout(embedding(inputs)).
    Parameters only; the arithmetic is dvae_bow_encoder_fwd / _bwd."""

    def __init__(self, vocab_size, emb_dim, emb_matrix=None, dropout_rate=0.5):
        super().__init__()
        self._device = torch.device("cpu")
        self.vocab_size, self.emb_dim = vocab_size, emb_dim
        if emb_matrix is not None:
            self.embedding = nn.Embedding.from_pretrained(torch.tensor(emb_matrix, dtype=torch.float32))
            self.embedding.weight.requires_grad = False
            self.vocab_size, self.emb_dim = emb_matrix.shape
        else:
            self.embedding = nn.Embedding(self.vocab_size, self.emb_dim)
        self.dropout = nn.Dropout(dropout_rate)
        self.dropout_rate = dropout_rate
        self.hidden_size, self.num_layers, self.num_directions = self.emb_dim, 1, 1      # "compatibility" (model.py:28-31)

    def forward(self, *a, **k):
        raise _lib.DvaeError("call VariationalSeq2Seq.forward(); sub-modules are parameter containers")


class VariationalEncoder(_DeviceMixin, nn.Module):
    """embedding -> dropout -> (bi)LSTM (vae/model.py:52-109); parameters only, see module doc."""

    def __init__(self, vocab_size, emb_dim, hidden_size, num_layers, dropout_rate=0.5, emb_matrix=None,
                 bidirectional=False):
        super().__init__()
        self._device = torch.device("cpu")
        self.vocab_size, self.emb_dim, self.hidden_size = vocab_size, emb_dim, hidden_size
        self.num_layers = num_layers
        self.num_directions = 2 if bidirectional is True else 1
        self.dropout_rate = dropout_rate
        if emb_matrix is not None:
            self.embedding = nn.Embedding.from_pretrained(torch.tensor(emb_matrix, dtype=torch.float32))
            self.embedding.weight.requires_grad = False
            self.vocab_size, self.emb_dim = emb_matrix.shape
        else:
            self.embedding = nn.Embedding(self.vocab_size, self.emb_dim)
        self.dropout = nn.Dropout(dropout_rate)
        self.recurrent = nn.LSTM(self.emb_dim, self.hidden_size, num_layers=self.num_layers,
                                 dropout=self.dropout_rate, batch_first=True, bidirectional=bidirectional)

    def init_hidden(self, batch_size):
        shape = (self.num_layers * self.num_directions, batch_size, self.hidden_size)
        return torch.zeros(*shape, device=self.device), torch.zeros(*shape, device=self.device)

    def forward(self, *a, **k):
        raise _lib.DvaeError("call VariationalSeq2Seq.encode(); sub-modules are parameter containers")


class VariationalDecoder(_DeviceMixin, nn.Module):
    """LSTM -> linear (vae/model.py:112-165); parameters only."""

    def __init__(self, vocab_size, emb_dim, hidden_size, num_layers, dropout_rate=0.5, emb_matrix=None):
        super().__init__()
        self._device = torch.device("cpu")
        self.vocab_size, self.emb_dim, self.hidden_size = vocab_size, emb_dim, hidden_size
        if num_layers == 1:          # vae/model.py:123-124
            num_layers = 2
        self.num_layers = num_layers
        self.dropout_rate = dropout_rate
        if emb_matrix is not None:
            self.embedding = nn.Embedding.from_pretrained(torch.tensor(emb_matrix, dtype=torch.float32))
            self.embedding.weight.requires_grad = False
            self.vocab_size, self.emb_dim = emb_matrix.shape
        else:
            self.embedding = nn.Embedding(self.vocab_size, self.emb_dim)
        self.dropout = nn.Dropout(self.dropout_rate)
        self.recurrent = nn.LSTM(self.emb_dim, self.hidden_size, num_layers=self.num_layers,
                                 dropout=self.dropout_rate, batch_first=True)
        self.linear = nn.Linear(self.hidden_size, self.vocab_size)

    def forward(self, *a, **k):
        raise _lib.DvaeError("call VariationalSeq2Seq.forward(); sub-modules are parameter containers")


class Discriminator(_DeviceMixin, nn.Module):
    """Linear probe on one latent space (vae/model.py:168-216)."""

    def __init__(self, name, latent_dim, output_dim):
        super().__init__()
        self._device = torch.device("cpu")
        self.name, self.latent_dim, self.output_dim = name, latent_dim, output_dim
        self.linear = nn.Linear(latent_dim, output_dim)
        assert self.output_dim > 0

    def predict(self, logits):
        # sigmoid(x) > 0.5  <=>  x > 0 ; softmax argmax == logits argmax  (vae/model.py:204-210)
        if logits.size(1) == 1:
            return (logits > 0).long().squeeze()
        return logits.argmax(-1).squeeze()

    def compute_loss(self, logits, targets):
        from . import losses
        return losses._single_dsc_loss(self, logits, targets)[0]

    def compute_accuracy(self, logits, targets):
        from . import losses
        return losses._single_dsc_loss(self, logits, targets)[1]


class AdversarialDiscriminator(Discriminator):
    """Tries to predict label `label_name` from latent space `latent_name` (vae/model.py:219-258).  The VAE is trained
    to maximise the entropy of its predictions; the adversary itself is trained, with its own Adam (lr 3e-4), on the
    DETACHED latent.  As in the reference, when `optimizer_step` runs the adversary's .grad already holds the
    entropy-term gradients from `total_loss.backward()` (run.py:254-260): both are applied."""

    def __init__(self, latent_name, label_name, latent_dim, output_dim):
        super().__init__(f"{latent_name}-{label_name}", latent_dim, output_dim)
        self.latent_name, self.label_name = latent_name, label_name
        self.optimizer = torch.optim.Adam(self.parameters(), lr=3e-4)
        self.detached_inputs = None

    def forward(self, inputs):
        from .functions import _SmallLinearFn
        self.detached_inputs = inputs.detach()
        return _SmallLinearFn.apply(inputs, self.linear.weight, self.linear.bias, 0)

    def compute_discriminator_loss(self, logits, targets):
        from . import losses
        from .functions import _SmallLinearFn
        detached_logits = _SmallLinearFn.apply(self.detached_inputs, self.linear.weight, self.linear.bias, 0)
        return losses._single_dsc_loss(self, detached_logits, targets)[0]

    def optimizer_step(self, dsc_loss):
        dsc_loss.backward(retain_graph=True)
        self.optimizer.step()
        self.optimizer.zero_grad()

    def compute_adversarial_loss(self, logits):
        from .functions import _EntropyLossFn
        if logits.dim() == 1:
            logits = logits.unsqueeze(1)
        elif logits.dim() > 2:
            raise ValueError(f"Got unexpected logits shape {logits.size()}")
        return _EntropyLossFn.apply(logits)        # = -H: minimising it maximises the entropy


class LatentParams(OrderedDict):
    """{space: Params(z, mu, logvar)} plus the fused per-space KL vector (autograd-connected)."""
    kl = None
    names = ()


class DscLogits(OrderedDict):
    """{label: logits [B,out]} plus the packed [B, sum(out)] tensor the fused loss kernel reads."""
    packed = None
    dims = None


class FusedLogits:
    """Stand-in for the reference's `decoder_logits` [B,T,V] tensor (vae/model.py:452-462).

    The B x T x V logits are never materialised on the hot path: `losses.reconstruction_loss`
    consumes (h_top, W, b) through the fused vocab-CE kernel.  `materialize()` builds the dense
    tensor on demand (debug / evaluation scripts that really want logits)."""

    def __init__(self, model, plan, h_top, B, T):
        self.model, self.plan, self.h_top = model, plan, h_top
        self.shape = (B, T, model.decoder.vocab_size)
        self.argmax_tokens = None

    def size(self, i=None):
        return self.shape if i is None else self.shape[i]

    def materialize(self):
        from .functions import materialize_logits
        return materialize_logits(self)

    def argmax(self, dim=-1):
        """Token-level reconstruction argmax [B,T] (position 0 is <SOS>)."""
        assert dim in (-1, 2)
        from .functions import vocab_argmax
        return vocab_argmax(self)


class VariationalSeq2Seq(_DeviceMixin, nn.Module):
    def __init__(self, encoder, decoder, discriminators, latent_dim, sos_token_idx, eos_token_idx,
                 adversarial_loss=False, mi_loss=False):
        super().__init__()
        self._device = torch.device("cpu")
        self.encoder, self.decoder, self.latent_dim = encoder, decoder, latent_dim
        self.discriminators = nn.ModuleDict()
        self.context2params = nn.ModuleDict()
        self.dsc_latent_dim = 0
        linear_insize = encoder.hidden_size * encoder.num_layers * encoder.num_directions
        for dsc in discriminators:                      # vae/model.py:287-294
            self.dsc_latent_dim += dsc.latent_dim
            self.discriminators[dsc.name] = dsc
            self.context2params[dsc.name] = nn.Linear(linear_insize, 2 * dsc.latent_dim)
        assert self.dsc_latent_dim <= self.latent_dim
        if self.dsc_latent_dim < self.latent_dim:       # vae/model.py:297-302
            leftover = self.latent_dim - self.dsc_latent_dim
            self.context2params["content"] = nn.Linear(linear_insize, 2 * leftover)
        self.adversarial_loss, self.mi_loss = adversarial_loss, mi_loss
        # same construction order as the reference (vae/model.py:304-317), so a seed gives the same initial weights
        self.adversaries = self._get_adversaries() if adversarial_loss is True else dict()
        self.mi_estimators = self._get_mi_estimators() if mi_loss is True else dict()
        self.z2hidden = nn.Linear(self.latent_dim, 2 * decoder.hidden_size * decoder.num_layers)
        self.sos_token_idx, self.eos_token_idx = sos_token_idx, eos_token_idx
        self._plans = {}
        self._P = None
        self._flat = None
        self._dims = None

    def _get_adversaries(self):
        """vae/model.py:323-335: one adversary per (latent space, label) pair with latent != label."""
        adversaries = nn.ModuleDict()
        for latent_name, layer in self.context2params.items():
            latent_size = layer.out_features // 2
            for label_name, dsc in self.discriminators.items():
                if latent_name == label_name:
                    continue
                adversaries[f"{latent_name}-{label_name}"] = AdversarialDiscriminator(latent_name, label_name, latent_size,
                                                                                   dsc.output_dim)
        return adversaries

    def _get_mi_estimators(self):
        """vae/model.py:337-355: one CLUB estimator per unordered pair of latent spaces (a plain dict, as in the
        reference: not part of state_dict / parameters())."""
        from . import losses
        mi_estimators, seen = dict(), set()
        for ni, li in self.context2params.items():
            for nj, lj in self.context2params.items():
                if ni == nj or (nj, ni) in seen:
                    continue
                seen.add((ni, nj))
                si, sj = li.out_features // 2, lj.out_features // 2
                mi_estimators[f"{ni}-{nj}"] = losses.CLUB(si, sj, max([si, sj, 5]))
        return mi_estimators

    # ---- parameter fusion: one flat fp32 buffer, nn.Parameters are views into it ----------------
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        for est in self.mi_estimators.values():       # plain dict: nn.Module._apply does not reach them
            est._apply(fn, *a, **k)
        self._fuse_parameters()
        return out

    def trainable_parameters(self):
        return [p for (n, p) in self.named_parameters() if p.requires_grad is True and not n.startswith("adversaries")]

    def flat_order(self):
        """Names of the trainable parameters in flat-buffer order (grouped so that the fused-head
        kernel sees the per-space context2params / discriminator tensors as single matrices)."""
        named = OrderedDict((n, p) for n, p in self.named_parameters()
                            if p.requires_grad and not n.startswith("adversaries"))     # adversaries own their optimizers
        spaces = list(self.context2params.keys())
        c2p_w = [f"context2params.{s}.weight" for s in spaces]
        c2p_b = [f"context2params.{s}.bias" for s in spaces]
        dsc = [s for s in spaces if s in self.discriminators]
        dsc_w = [f"discriminators.{s}.linear.weight" for s in dsc]
        dsc_b = [f"discriminators.{s}.linear.bias" for s in dsc]
        grouped = set(c2p_w + c2p_b + dsc_w + dsc_b)
        rest = [n for n in named if n not in grouped]
        return rest, [c2p_w, c2p_b, dsc_w, dsc_b]

    def _fuse_parameters(self):
        named = OrderedDict(self.named_parameters())
        if not named:
            return
        dev = next(iter(named.values())).device
        rest, groups = self.flat_order()
        layout, off = OrderedDict(), 0
        for n in rest:
            layout[n] = off
            off += (named[n].numel() + 3) // 4 * 4
        group_off = []
        for g in groups:
            group_off.append(off)
            for n in g:
                layout[n] = off
                off += named[n].numel()
            off = (off + 3) // 4 * 4
        flat = torch.zeros(max(off, 4), device=dev, dtype=torch.float32)
        for n, o in layout.items():
            p = named[n]
            view = flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self._flat, self._layout, self._flat_numel = flat, layout, off
        P = {n: p.data for n, p in named.items()}
        spaces = list(self.context2params.keys())
        C = self.encoder.hidden_size * self.encoder.num_layers * self.encoder.num_directions
        Z2 = sum(self.context2params[s].out_features for s in spaces)
        P["_c2p.weight"] = flat[group_off[0]:group_off[0] + Z2 * C].view(Z2, C)
        P["_c2p.bias"] = flat[group_off[1]:group_off[1] + Z2]
        n_w = sum(named[n].numel() for n in groups[2])
        n_b = sum(named[n].numel() for n in groups[3])
        if n_w:
            P["_dsc.weight"] = flat[group_off[2]:group_off[2] + n_w]
            P["_dsc.bias"] = flat[group_off[3]:group_off[3] + n_b]
        self._P, self._group_off, self._groups = P, group_off, groups
        self._plans = {}
        self._dims = Dims(self)

    def grad_views(self, flat_grad):
        """{name: view} into a flat gradient buffer laid out like the parameters."""
        named = OrderedDict(self.named_parameters())
        G = {n: flat_grad[o:o + named[n].numel()].view(named[n].shape) for n, o in self._layout.items()}
        go = self._group_off
        G["_c2p.weight"] = flat_grad[go[0]:go[0] + self._P["_c2p.weight"].numel()]
        G["_c2p.bias"] = flat_grad[go[1]:go[1] + self._P["_c2p.bias"].numel()]
        if "_dsc.weight" in self._P:
            G["_dsc.weight"] = flat_grad[go[2]:go[2] + self._P["_dsc.weight"].numel()]
            G["_dsc.bias"] = flat_grad[go[3]:go[3] + self._P["_dsc.bias"].numel()]
        return G

    def _require_cuda(self):
        if self._flat is None:
            self._fuse_parameters()
        if self._flat.device.type != "cuda":
            raise _lib.DvaeError("the B200 path needs the model on a CUDA device (build_vae(..., device=cuda)); "
                                 "there is no CPU fallback")

    def get_plan(self, B, T, need_grad):
        self._require_cuda()
        pool = self._plans.setdefault((B, T), [])
        for pl in pool:
            if not pl.busy:
                break
        else:
            pl = StepPlan(self, B, T, self._flat.device)
            pool.append(pl)
        pl.busy = bool(need_grad)
        return pl

    # ---- reference surface ---------------------------------------------------------------------
    def _draw_eps(self, B, device):
        """Reparameterisation noise in the reference's draw order: per space, two `randn` calls in
        train mode (the second is the one used), one in eval mode (vae/model.py:391-395)."""
        eps = []
        for name, layer in self.context2params.items():
            zs = layer.out_features // 2
            if self.training:
                torch.randn(B, zs, device=device)
            eps.append(torch.randn(B, zs, device=device))
        return torch.cat(eps, dim=1).contiguous()

    def encode(self, inputs, lengths):
        from .functions import run_encoder
        return run_encoder(self, inputs, lengths)

    def compute_latent_params(self, context, eps=None):
        from .functions import run_heads
        return run_heads(self, context, eps)[0]

    def compute_hidden(self, z, batch_size):
        from .functions import run_z2hidden
        return run_z2hidden(self, z)

    def forward(self, inputs, lengths, teacher_forcing_prob=0.5, eps=None):
        """Same contract as vae/model.py:413-482.  `eps` ([B,Z], optional extension) replays given
        reparameterisation noise instead of drawing it."""
        from .functions import run_forward
        # one coin per decoding step, shared by the batch, from Python's RNG (vae/model.py:463)
        T = inputs.size(-1)
        coins = [random.random() < teacher_forcing_prob for _ in range(1, T)]
        return run_forward(self, inputs, lengths, coins, eps)

    def sample(self, z, max_length=30):
        from .functions import run_sample
        return run_sample(self, z, max_length)


def build_vae(params, vocab_size, emb_matrix, label_dims, device, sos_token_idx, eos_token_idx):
    """vae/model.py:515-559 -- same arguments, same construction order (so the same seed gives the
    same initial weights as the reference)."""
    if params["bow_encoder"] is True:
        encoder = BOWEncoder(vocab_size, params["embedding_dim"], emb_matrix=emb_matrix,
                             dropout_rate=params["encoder_dropout"])
    else:
        encoder = VariationalEncoder(vocab_size, params["embedding_dim"], params["hidden_dim"], params["num_rnn_layers"],
                                     dropout_rate=params["encoder_dropout"], emb_matrix=emb_matrix,
                                     bidirectional=params["bidirectional_encoder"])
    encoder.set_device(device)
    decoder = VariationalDecoder(vocab_size, params["embedding_dim"], params["hidden_dim"], params["num_rnn_layers"],
                                 dropout_rate=params["decoder_dropout"], emb_matrix=emb_matrix)
    decoder.set_device(device)
    discriminators = []
    for (name, outdim) in label_dims.items():
        if name not in params["latent_dims"]:
            continue
        dsc = Discriminator(name, params["latent_dims"][name], outdim)
        dsc.set_device(device)
        discriminators.append(dsc)
    vae = VariationalSeq2Seq(encoder, decoder, discriminators, params["latent_dims"]["total"], sos_token_idx,
                             eos_token_idx, adversarial_loss=params["adversarial_loss"], mi_loss=params["mi_loss"])
    vae.set_device(device)
    return vae
