"""Graph-captured data-parallel training engine: the reference's train-step body
(`run.py:217-262`: forward -> compute_all_losses -> backward -> clip_grad_norm_(5.0) -> Adam.step
-> zero_grad) as one fixed sequence of C-ABI kernel launches over pre-allocated buffers, replayed
as CUDA graphs, with an NCCL all-reduce of the flat gradient buffer between backward and the
optimiser tail when world_size > 1.

Same kernels and the same `StepPlan` as the drop-in autograd path (`functions.py`); what the
engine removes is per-op host work (autograd bookkeeping, ctypes marshalling, allocator traffic).
"""
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib import ptr, check
from .plan import StepPlan
from .losses import get_cyclic_kl_weight


def d_bow(plan):
    return bool(plan.d.bow)


class TrainEngine:
    def __init__(self, model, params, B, T, lr=None, total_steps=None, use_graph=True, max_norm=5.0,
                 process_group=None, seed=None):
        model._require_cuda()
        if len(model.adversaries) or len(model.mi_estimators):
            raise NotImplementedError("TrainEngine captures the ELBO + discriminator step; models with adversarial_loss / mi_loss "
                                      "train through the drop-in path (forward / compute_all_losses / backward, run.py:217-276)")
        self.model, self.params, self.B, self.T = model, params, B, T
        self.lib = _lib.load()
        self.device = model._flat.device
        self.plan = StepPlan(model, B, T, self.device)
        self.d = d = self.plan.d
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.pg = process_group
        self.max_norm = max_norm
        self.lr = float(params["learn_rate"] if lr is None else lr)
        self.total_steps = total_steps
        self.lambdas = params["lambdas"]
        self.n = model._flat_numel
        f32 = dict(device=self.device, dtype=torch.float32)
        self.flat = model._flat[:self.n]
        self.grad = torch.zeros(self.n, **f32)
        self.m = torch.zeros(self.n, **f32)
        self.v = torch.zeros(self.n, **f32)
        self.G = model.grad_views(self.grad)
        named = dict(model.named_parameters())
        for n in model._layout:
            named[n].grad = self.G[n]
        # per-step scalars live in device memory so one captured graph serves every step
        self.n_dsc = max(sum(1 for o in d.dsc_out if o > 0), 1)
        self.hyper = torch.zeros(8, **f32)           # lr, beta1, beta2, eps, step
        self.kl_w = torch.zeros(max(d.S, 1), **f32)
        self.sumsq = torch.zeros(1, **f32)
        self.red_ws = torch.zeros(1032, **f32)
        self.inputs = torch.zeros(B, T, device=self.device, dtype=torch.int64)
        self.lengths = torch.zeros(B, device=self.device, dtype=torch.int64)
        self.labels = torch.zeros(self.n_dsc, B, **f32)
        # pinned staging for the end-to-end (host buffers in, loss out) entry point
        self.h_inputs = torch.zeros(B, T, dtype=torch.int64).pin_memory()
        self.h_lengths = torch.zeros(B, dtype=torch.int64).pin_memory()
        self.h_labels = torch.zeros(self.n_dsc, B, dtype=torch.float32).pin_memory()
        self.h_scal = torch.zeros(8 + max(d.S, 1), dtype=torch.float32).pin_memory()
        self.h_seed = torch.zeros(1, dtype=torch.int64).pin_memory()
        self.h_out = torch.zeros(self.plan.out.numel(), dtype=torch.float32).pin_memory()
        self.d_scal = torch.zeros(8 + max(d.S, 1), **f32)
        self.step_idx = 0
        self._gen = torch.Generator().manual_seed(int(params.get("random_seed", 10)) if seed is None else seed)
        self.use_graph = use_graph
        self._graphs = None
        self._buckets = None
        self._comm = None
        self._side = None
        self.label_names = [n for n, o in zip(d.space_names, d.dsc_out) if o > 0]

    # ---- the kernel sequences ------------------------------------------------------------------
    def _fwd_bwd(self, part=None):
        """part = None: the whole forward + backward; 1: forward, vocabulary backward and decoder backward (every
        decoder.* gradient final); 2: heads and encoder backward.  The split lets the all-reduce of the decoder
        gradients (half of the parameters) run under part 2 when training data-parallel."""
        pl, P, G, m = self.plan, self.model._P, self.G, self.model
        st = _lib.stream_ptr()
        if part in (None, 1):
            # unpack the per-step scalar block (device-to-device, inside the graph)
            self.hyper[:5].copy_(self.d_scal[:5])
            self.kl_w.copy_(self.d_scal[8:8 + self.kl_w.numel()])
            pl.randn_eps()
            # the decoder's input embeddings, its layer-0 input projection and the W_out operand planes do not depend on
            # the encoder (teacher forcing): they run on a side stream under the encoder's recurrences (64 of 148 SMs)
            cur = torch.cuda.current_stream()
            hoist = os.environ.get("DVAE_HOIST", "1") != "0" and not d_bow(pl)
            fork_prep = None
            if hoist:
                if self._side is None:
                    self._side = torch.cuda.Stream(device=self.device)

                def fork_prep():      # forked behind the encoder's layer-0 projection GEMMs, which fill the machine themselves
                    self._side.wait_stream(cur)
                    with torch.cuda.stream(self._side):
                        pl.decode_prepare(P, self.inputs, m.sos_token_idx, True)
            pl.encode(P, self.inputs, self.lengths, True, after_l0_proj=fork_prep)
            pl.heads(P, pl.ctx, pl.eps, self.labels, self.kl_w)
            if hoist:
                cur.wait_stream(self._side)
            h_top = pl.decode_forced(P, self.inputs, m.sos_token_idx, True)
            pl.vocab_ce(P, h_top, self.inputs, self.lengths)
            g_top = pl.vocab_ce_bwd(P, G, h_top, self.inputs, self.lengths, None)
            # weight-gradient GEMMs of each layer keep running on side streams while the next layer's recurrence starts
            check(self.lib.dvae_defer_joins(1), "dvae_defer_joins")
            try:
                pl.decode_bwd(P, G, g_top, emb_grad="decoder.embedding.weight" in m._layout,
                              aux_stream=self._side if hoist else None)
            finally:
                if part == 1:
                    check(self.lib.dvae_join_side_streams(st), "dvae_join_side_streams")
                    if hoist:
                        cur.wait_stream(self._side)
            self._aux_pending = hoist and part is None
        if part in (None, 2):
            check(self.lib.dvae_defer_joins(1), "dvae_defer_joins")
            try:
                g_ctx = pl.heads_bwd(P, G, pl.ctx, pl.eps, self.labels, self.kl_w, pl.g_hid)
                if getattr(self, "_aux_pending", False):      # the decoder's embedding-gradient scatter still reads g_dx[0]
                    torch.cuda.current_stream().wait_stream(self._side)
                    self._aux_pending = False
                pl.encode_bwd(P, G, self.inputs, self.lengths, g_ctx, emb_grad="encoder.embedding.weight" in m._layout)
            finally:
                check(self.lib.dvae_join_side_streams(st), "dvae_join_side_streams")

    def _optim(self):
        st = _lib.stream_ptr()
        check(self.lib.dvae_grad_sumsq(ptr(self.grad), self.n, ptr(self.sumsq), ptr(self.red_ws), st), "dvae_grad_sumsq")
        check(self.lib.dvae_clip_adam(ptr(self.flat), ptr(self.grad), ptr(self.m), ptr(self.v), self.n, ptr(self.sumsq),
                                      self.max_norm, 1.0 / self.world, ptr(self.hyper), 1, st), "dvae_clip_adam")

    def _grad_buckets(self):
        from .dist import grad_buckets
        return grad_buckets(self.model, self.grad)

    def _run(self):
        if self.world == 1:
            if not self.use_graph:
                self._fwd_bwd()
                self._optim()
                return
            if self._graphs is None:
                self._capture()
            self._graphs[0].replay()          # one graph: forward, backward, clip + Adam
            return
        # data parallel: all-reduce the decoder gradients on a side stream while heads + encoder backward run
        if self._buckets is None:
            self._buckets = self._grad_buckets()
            self._comm = torch.cuda.Stream(device=self.device)
        dec_bucket, rest = self._buckets
        cur = torch.cuda.current_stream()
        if self.use_graph and self._graphs is None:
            self._capture()
        if self.use_graph:
            self._graphs[0].replay()
        else:
            self._fwd_bwd(1)
        overlap = os.environ.get("DVAE_DP_OVERLAP", "1") != "0"
        if overlap:
            self._comm.wait_stream(cur)
            with torch.cuda.stream(self._comm):
                dist.all_reduce(dec_bucket, op=dist.ReduceOp.SUM, group=self.pg)
        if self.use_graph:
            self._graphs[1].replay()
        else:
            self._fwd_bwd(2)
        if overlap:
            for b in rest:
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=self.pg)
            cur.wait_stream(self._comm)
        else:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=self.pg)
        if self.use_graph:
            self._graphs[2].replay()
        else:
            self._optim()

    def _capture(self):
        # warm up eagerly on a side stream (lazy module loading, cudaFuncSetAttribute, allocator), then capture
        snap = (self.flat.clone(), self.m.clone(), self.v.clone())
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._fwd_bwd()
                self._optim()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graphs = []
        if self.world == 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self._fwd_bwd(None)
                self._optim()
            graphs.append(g)
        else:                 # the all-reduces sit between the graphs (NCCL on its own stream)
            for part in (1, 2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    self._fwd_bwd(part)
                graphs.append(g)
            gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gb, stream=s):
                self._optim()
            graphs.append(gb)
        torch.cuda.synchronize()
        # undo the warm-up updates so that capture leaves the training state untouched
        self.flat.copy_(snap[0]); self.m.copy_(snap[1]); self.v.copy_(snap[2])
        self.grad.zero_()
        self._graphs = tuple(graphs)

    # ---- per-step host scalars -----------------------------------------------------------------
    def _fill_scalars(self):
        step = self.step_idx
        h = self.h_scal
        h[0], h[1], h[2], h[3], h[4] = self.lr, 0.9, 0.999, 1e-8, float(step + 1)
        for i, n in enumerate(self.d.space_names):
            w = self.lambdas[n] if n in self.lambdas else self.lambdas["default"]
            if w == "cyclic":
                w = get_cyclic_kl_weight(step, self.total_steps if self.total_steps else 1)
            h[8 + i] = float(w)
        self.h_seed[0] = int(torch.randint(0, 2 ** 62, (1,), generator=self._gen))

    # ---- public entry points -------------------------------------------------------------------
    def step_resident(self, inputs_dev, lengths_dev, labels_dev):
        """One train step on a batch already in HBM; returns the device result block
        (plan.out: [0] weighted KL, [1] KL, [2] dsc loss, [3..] per space, [27] reconstruction)."""
        self._fill_scalars()
        self.d_scal.copy_(self.h_scal, non_blocking=True)
        self.plan.seed_dev.copy_(self.h_seed, non_blocking=True)
        self.inputs.copy_(inputs_dev, non_blocking=True)
        self.lengths.copy_(lengths_dev, non_blocking=True)
        self.labels.copy_(labels_dev, non_blocking=True)
        self._run()
        self.step_idx += 1
        return self.plan.out

    def step_host(self, inputs, lengths, labels):
        """End-to-end step: HOST tensors in (inputs [B,T] int64, lengths [B] int64, labels {name: [B,1]}),
        python floats out.  Copies host->pinned->device, runs the step, reads the loss block back."""
        self.h_inputs.copy_(inputs)
        self.h_lengths.copy_(lengths)
        for i, n in enumerate(self.label_names):
            self.h_labels[i].copy_(labels[n].reshape(-1))
        self._fill_scalars()
        self.d_scal.copy_(self.h_scal, non_blocking=True)
        self.plan.seed_dev.copy_(self.h_seed, non_blocking=True)
        self.inputs.copy_(self.h_inputs, non_blocking=True)
        self.lengths.copy_(self.h_lengths, non_blocking=True)
        self.labels.copy_(self.h_labels, non_blocking=True)
        self._run()
        self.h_out.copy_(self.plan.out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.step_idx += 1
        return self.losses_from(self.h_out)

    def losses_from(self, out):
        S, NS = self.d.S, _lib.HEADS_NSCALARS
        o = out.tolist()
        L = {"reconstruction_loss": o[NS], "total_weighted_kl": o[0], "total_kl": o[1], "total_dsc_loss": o[2],
             "idv_kls": {n: o[3 + i] for i, n in enumerate(self.d.space_names)},
             "idv_dsc_losses": {n: o[3 + S + i] for i, n in enumerate(self.d.space_names) if self.d.dsc_out[i] > 0},
             "idv_dsc_accs": {n: o[3 + 2 * S + i] for i, n in enumerate(self.d.space_names) if self.d.dsc_out[i] > 0}}
        L["total_loss"] = o[NS] + o[0] + o[2]
        return L

    @property
    def h2d_bytes_per_step(self):
        return (self.h_inputs.numel() + self.h_lengths.numel() + self.h_seed.numel()) * 8 + \
            (self.h_labels.numel() + self.h_scal.numel()) * 4

    @property
    def d2h_bytes_per_step(self):
        return self.h_out.numel() * 4
