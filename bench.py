#!/usr/bin/env python
"""Headline benchmark: device-timed training tokens/sec of the disentangled sentence VAE step
(forward + all losses + backward + [NCCL all-reduce] + clip + Adam) at BASELINE.json configs[1]
(the sfu_amazon_100k reproduction shape) on N B200s, plus roofline / CPU-baseline / end-to-end legs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
Synthetic data of the reference's shape, reference initialisation under seed 10.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAD, UNK, SOS, EOS = 0, 1, 2, 3

# BASELINE.json configs[1]: reproduction_configs/sfu_amazon_100k/vae_elbo.json with "default": "cyclic"
CFG2 = {"name": "bench/sfu_amazon_100k", "random_seed": 10, "data_dir": "", "combined_dataset": True,
        "dataset_minibatch_ratios": {"sfu": 0.5, "amazon": 0.5}, "checkpoint_dir": "", "glove_path": "",
        "num_train_examples": -1, "lowercase": True, "reverse_input": False, "embedding_dim": 256,
        "hidden_dim": 256, "num_rnn_layers": 2, "bidirectional_encoder": True, "bow_encoder": False,
        "latent_dims": {"total": 64, "polarity": 1, "uncertainty": 1}, "epochs": 20, "batch_size": 128,
        "learn_rate": 3e-4, "encoder_dropout": 0.5, "decoder_dropout": 0.5, "teacher_forcing_prob": 1.0,
        "lambdas": {"default": "cyclic", "polarity": 0.005, "uncertainty": 0.005}, "adversarial_loss": False,
        "mi_loss": False, "train": True, "validate": False, "test": False}
VOCAB, SEQ_T = 10000, 22
TOTAL_STEPS = 20 * 1563          # epochs * len(dataloader) of the reproduction run
LABELS = {"uncertainty": 1, "polarity": 1}
WORKLOAD = ("cfg2 sfu_amazon_100k reproduction shape: per-GPU batch 128, T=22 (SFU length histogram: min 3, mean ~10, "
            "max 22), V=10000, E=H=256, 2-layer bi-LSTM encoder, 2-layer decoder, Z=64 (uncertainty 1, polarity 1, "
            "content 62), dropout 0.5, teacher forcing 1.0, cyclic KL")


def synth_batch(rng, B, T=None, V=None, uniform_lengths=None):
    """rows [SOS, w..., EOS, PAD...], Zipf(1.0) body tokens, SFU-like lengths (or U{lo..hi}), 10% labels."""
    T = SEQ_T if T is None else T
    V = VOCAB if V is None else V
    if uniform_lengths is not None:
        lengths = rng.integers(uniform_lengths[0], uniform_lengths[1] + 1, B).astype(np.int64)
    else:
        lengths = np.clip(np.round(rng.gamma(9.7, 1.06, B)), 3, T).astype(np.int64)
    lengths[0] = T
    ranks = np.arange(4, V)
    pz = 1.0 / (ranks - 3.0)
    pz /= pz.sum()
    X = np.zeros((B, T), np.int64)
    for b in range(B):
        n = lengths[b]
        X[b, 0], X[b, n - 1] = SOS, EOS
        X[b, 1:n - 1] = rng.choice(ranks, size=n - 2, p=pz)
    Y = np.stack([(rng.random(B) < 0.1).astype(np.float32) for _ in LABELS])
    return X, lengths, Y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.path = f"/tmp/dvae_clocks_{os.getpid()}.csv"
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "src": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------------------------------------
# CPU legs: the oracle port timed on the host cores (reported baseline only)
# ------------------------------------------------------------------------------------------------
def cpu_port_steps(n_steps, warmup, B=128, budget_s=150.0):
    """Times `oracle/dvae_oracle.py` (numpy, float32, BLAS threads = all host cores) on the cfg-2
    workload: forward + losses + backward + clip + Adam.  Returns (tokens/s, s/step, sample description)."""
    from oracle import dvae_oracle as O
    dvae = importlib.import_module("disentanglement-vae_b200")
    dvae.set_seed(10)
    vae = dvae.build_vae(CFG2, VOCAB, None, LABELS, torch.device("cpu"), SOS, EOS)     # weights only; never run
    sd = O.cast_state_dict({k: v.detach().numpy() for k, v in vae.state_dict().items()}, np.float32)
    spec = O.ModelSpec(sd, list(vae.context2params.keys()), SOS, EOS)
    rng = np.random.default_rng(10)
    m = {k: np.zeros_like(v) for k, v in sd.items()}
    v_ = {k: np.zeros_like(v) for k, v in sd.items()}
    names = list(LABELS)

    def one(i, Bs):
        X, lengths, Y = synth_batch(rng, Bs)
        eps = {n: rng.standard_normal((Bs, zs)).astype(np.float32) for n, zs in zip(spec.space_names, spec.space_dims)}
        masks_e = [(rng.random((SEQ_T, Bs, w)) >= 0.5).astype(np.float32) * 2 for w in (256, 512)]
        masks_d = [(rng.random((SEQ_T - 1, Bs, w)) >= 0.5).astype(np.float32) * 2 for w in (256, 256)]
        klw = {"default": O.cyclic_kl_weight(i, TOTAL_STEPS), "polarity": 0.005, "uncertainty": 0.005}
        t0 = time.perf_counter()
        fw = O.model_forward(sd, spec, X, lengths, eps, labels={n: Y[j].reshape(-1, 1) for j, n in enumerate(names)},
                             kl_weights=klw, enc_masks=masks_e, dec_masks=masks_d)
        g = O.model_backward(sd, spec, fw)
        O.clip_and_adam(sd, g, m, v_, i + 1, CFG2["learn_rate"])
        return time.perf_counter() - t0, int(lengths.sum())

    t_probe, _ = one(0, B)
    Bs = B
    total = n_steps + max(warmup - 1, 0)
    if t_probe * total > budget_s:                       # bounded sample: shrink the batch, keep the shape
        Bs = max(8, int(B * budget_s / (t_probe * total)) // 8 * 8)
    for i in range(max(warmup - 1, 0)):
        one(i + 1, Bs)
    ts, toks = 0.0, 0
    for i in range(n_steps):
        dt, nt = one(warmup + i, Bs)
        ts += dt
        toks += nt
    sample = f"{n_steps} train steps of the cfg-2 workload at batch {Bs} (numpy float32 port of vae/model.py + vae/losses.py + run.py:254-262)"
    return toks / ts, ts / n_steps, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    tps, sps, sample = cpu_port_steps(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "train_tokens_per_sec", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# Secondary workloads (BASELINE.json configs[3] / configs[4]); the headline line is always cfg 2.
# ------------------------------------------------------------------------------------------------
def select_workload(name):
    """Rebinds the module-level workload description.  cfg4 = scaled decoder (H 1024, V 50k, T 64: stresses the fused
    vocab-CE kernels; SURVEY.md 8 states B = 128, lengths U{16..64})."""
    global CFG2, VOCAB, SEQ_T, WORKLOAD, TOTAL_STEPS
    if name == "cfg4":
        CFG2 = dict(CFG2, name="bench/scaled_decoder", hidden_dim=1024)
        VOCAB, SEQ_T = 50000, 64
        WORKLOAD = ("cfg4 scaled decoder: per-GPU batch 128, T=64 (lengths U{16..64}), V=50000, E=256, H=1024, 2-layer bi-LSTM "
                    "encoder, 2-layer decoder, Z=64, dropout 0.5, teacher forcing 1.0, cyclic KL")
        return (16, 64)
    return None


def run_inference_workload(args):
    """cfg5: consistency-evaluation shape (scripts/evaluation/consistency.py:163-205): for each batch of 1024 sentences,
    R = 30 resamples of [forward(x, tf=0) -> sampled reconstruction x_hat -> on-device length recount -> forward(x_hat,
    tf=0)], model in train mode (fresh latents and dropout each time), no gradients.  One 'step' = one resample pair."""
    import __graft_entry__ as ge
    dvae = ge.build()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B, T, R = 1024, SEQ_T, 30
    dvae.set_seed(10)
    vae = dvae.build_vae(CFG2, VOCAB, None, LABELS, dev, SOS, EOS)
    vae.train()
    rng = np.random.default_rng(5)
    X, L, _ = synth_batch(rng, B)
    Xd, Ld = torch.from_numpy(X).to(dev), torch.from_numpy(L).to(dev)

    def pair():
        with torch.no_grad():
            out = vae(Xd, Ld, teacher_forcing_prob=0.0)
            xh = out["token_predictions"]
            lh = xh.size(1) - ((xh == EOS) | (xh == PAD)).sum(1)          # consistency.py:186-190
            lh = lh.clamp_(min=1)
            out2 = vae(xh, lh, teacher_forcing_prob=0.0)
            return out2["dsc_logits"]

    n0 = dvae.launch_count()
    for _ in range(max(args.warmup, 3)):
        pair()
    torch.cuda.synchronize()
    launches = (dvae.launch_count() - n0) // max(args.warmup, 3)
    steps = args.steps if args.steps != 50 else R
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(0)
    a.record()
    for _ in range(steps):
        pair()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    clk = clocks.stop()
    toks = float(L.sum()) * 2 * steps
    line = {"metric": "inference_tokens_per_sec", "value": toks / (ms * 1e-3), "unit": "tokens/s", "n_gpus": 1, "steps": steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg5 inference: batch 1024, T=22, V=10000, E=H=256; per step = forward(x, tf=0) with sampled "
                                   "decoding + re-encode of the sampled sentences (consistency.py loop body), 30 resamples per batch",
                       "global_batch": B, "seq_len": T, "tokens": "valid input tokens x 2 forwards"},
            "gpu_launches": int(launches * steps), "launches_per_step": int(launches), "clocks": clk}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=128, help="per-GPU batch")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4", "cfg5"],
                    help="cfg2 = headline (BASELINE.json configs[1]); cfg4 / cfg5 = secondary configs, recorded under profiles/")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "cfg5":
        return run_inference_workload(args)
    uniform_lengths = select_workload(args.workload)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at first use: keep stdout clean for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    import __graft_entry__ as ge
    dvae = ge.build()
    from importlib import import_module
    engine_mod = import_module("disentanglement-vae_b200.engine")

    B, T = args.batch, SEQ_T
    dvae.set_seed(10)                                     # same initial weights on every rank
    vae = dvae.build_vae(CFG2, VOCAB, None, LABELS, dev, SOS, EOS)
    vae.train()
    eng = engine_mod.TrainEngine(vae, CFG2, B, T, total_steps=TOTAL_STEPS, use_graph=not args.no_graph, seed=10 + rank)
    rng = np.random.default_rng(1000 + rank)             # each rank draws its own shard of the global batch
    pool = [synth_batch(rng, B, uniform_lengths=uniform_lengths) for _ in range(8)]
    dpool = [(torch.from_numpy(X).to(dev), torch.from_numpy(L).to(dev), torch.from_numpy(Y).to(dev)) for X, L, Y in pool]
    hpool = [(torch.from_numpy(X), torch.from_numpy(L), {n: torch.from_numpy(Y[j]) for j, n in enumerate(LABELS)}) for X, L, Y in pool]
    tokens = [int(L.sum()) for _, L, _ in pool]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (includes graph capture) ----
    n0 = dvae.launch_count()
    eng.use_graph = False
    eng.step_resident(*dpool[0])                         # one eager step: counts this library's launches per step
    torch.cuda.synchronize()
    launches_per_step = dvae.launch_count() - n0
    eng.use_graph = not args.no_graph
    for i in range(max(args.warmup, 3)):
        out = eng.step_resident(*dpool[i % len(dpool)])
    barrier()
    loss_warm = eng.losses_from(out.cpu())["total_loss"]

    # ---- timed region: K steps, device-timed per step, L2 flushed between steps ----
    clocks = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    tok_sum = 0
    for i in range(args.steps):
        flush.zero_()
        j = (i + 1) % len(dpool)
        ev[i][0].record()
        out = eng.step_resident(*dpool[j])
        ev[i][1].record()
        tok_sum += tokens[j]
    barrier()
    clk = clocks.stop() if clocks else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    loss_last = eng.losses_from(out.cpu())["total_loss"]

    # ---- end-to-end leg: host buffers in, losses out, through the public engine API ----
    barrier()
    t0 = time.perf_counter()
    e2e_tok = 0
    for i in range(args.steps):
        j = i % len(hpool)
        L = eng.step_host(*hpool[j])
        e2e_tok += tokens[j]
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- dominant-kernel roofline: vocab-CE forward kernel timed alone with CUDA events ----
    pl, P = eng.plan, vae._P
    d = pl.d
    N = pl.N
    evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    from importlib import import_module as _im
    _L = _im("disentanglement-vae_b200._lib")
    pl.vocab_ce(P, pl.d_hs[-1], eng.inputs, eng.lengths)            # leaves the fp16 operand planes in the workspace
    call_ms = []
    for a, b in evk:                                                 # the whole C-ABI call: split + kernel + finalize + loss
        flush.zero_()
        a.record()
        pl.vocab_ce(P, pl.d_hs[-1], eng.inputs, eng.lengths)
        b.record()
    torch.cuda.synchronize()
    call_ms = float(np.median([a.elapsed_time(b) for a, b in evk]))
    for a, b in evk:                                                 # the kernel alone (one launch between the events)
        flush.zero_()
        a.record()
        _L.check(pl.lib.dvae_vocab_ce_partials(_L.ptr(pl.d_hs[-1]), d.Hd, pl.T1, pl.B, d.Hd, d.V, _L.ptr(P["decoder.linear.weight"]),
                                                _L.ptr(P["decoder.linear.bias"]), _L.ptr(eng.inputs), eng.inputs.stride(0),
                                                _L.ptr(eng.lengths), d.sos, _L.ptr(pl.ce_ws), _L.stream_ptr()), "dvae_vocab_ce_partials")
        b.record()
    torch.cuda.synchronize()
    k_ms = float(np.median([a.elapsed_time(b) for a, b in evk]))

    # ---- secondary roofline: the fused clip + Adam + zero_grad tail (HBM-bound), timed alone ----
    eva = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for a, b in eva:
        flush.zero_()
        a.record()
        eng._optim()
        b.record()
    torch.cuda.synchronize()
    adam_ms = float(np.median([a.elapsed_time(b) for a, b in eva]))

    # ---- K2 latent heads (one launch) and K1 encoder forward (embedding, 4 input GEMMs, 2 recurrence kernels), timed alone ----
    def _timed(fn, n=10):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for a, b in evs:
            flush.zero_()
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in evs]))
    heads_ms = _timed(lambda: pl.heads(P, pl.ctx, pl.eps, eng.labels, eng.kl_w))
    enc_ms = _timed(lambda: pl.encode(P, eng.inputs, eng.lengths, True))

    stats = torch.tensor([dev_ms, e2e_s, float(tok_sum), float(e2e_tok)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s = float(mx[0]), float(mx[1])
        tok_sum, e2e_tok = float(sm[2]), float(sm[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    flops = 2.0 * N * d.H * d.V
    tf = flops / (k_ms * 1e-3) / 1e12
    alg_bytes = 4.0 * (N * d.H + d.V * d.H + d.V + 2 * N)
    traffic = None
    try:                       # dram__bytes_read + write of this kernel from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")) as f:
            t = json.load(f)["vocab_ce_fwd"]
        traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    gemm_impl = os.environ.get("DVAE_GEMM_IMPL", "")
    kname = ("tc_gemm_kernel mode 1 (TMA + tcgen05.mma.kind::tf32, 3xTF32)" if gemm_impl.startswith("t")
             else "tc16_gemm_kernel mode 1 (pre-split fp16 hi/lo operand planes by bulk copy, A rows stationary in SMEM, tcgen05.mma.kind::f16, two TMEM accumulators, 16 epilogue warps)")
    n_par = eng.n
    Zt = d.Z
    heads_bytes = 4.0 * (B * d.C + d.C * 2 * Zt + 2 * Zt + B * Zt + 3 * B * Zt + Zt * d.H2L + B * d.H2L)
    Dn = 2 if CFG2.get("bidirectional_encoder", True) else 1
    enc_flops = float(B * T) * (Dn * 8 * d.H * (d.E + d.H) + Dn * 8 * d.H * (Dn * d.H + d.H))                 # SURVEY 8a, per position
    adam_bytes = 36.0 * n_par                   # sumsq reads g; clip+Adam reads p, g, m, v and writes p, m, v and the zeroed g
    roof = {"kernel": kname + " = vocab-CE forward: vocabulary projection + online log-softmax / arg-max / NLL epilogue from TMEM; "
                              "the [N,V] logits are never written to HBM; ONE launch between the CUDA events (dvae_vocab_ce_partials)",
            "bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"],
            "peak_source": f"MEASURED_PEAKS.json bf16 burst ({peaks['src']}); fp32-grade emulation executes 3 tensor flops per algorithmic flop",
            "traffic": traffic, "traffic_source": "profiles/r1_ncu_traffic.json (ncu --set full, per launch)",
            "launch_ms": k_ms, "call_ms": call_ms, "call": "dvae_vocab_ce_fwd = operand split x2 + this kernel + finalize + loss",
            "tensor_flops_executed": 3 * flops, "algorithmic_flops": flops, "algorithmic_bytes": alg_bytes,
            "hbm_frac": alg_bytes / (k_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "secondary": [{"kernel": "sumsq_kernel + clip_adam_kernel (grad-norm clip 5.0 + Adam + zero_grad over the flat parameter buffer)",
                           "bound": "hbm", "achieved": adam_bytes / (adam_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": adam_bytes / (adam_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "launch_ms": adam_ms,
                           "algorithmic_bytes": adam_bytes},
                          {"kernel": "heads_fwd_kernel (context2params, reparameterisation, KL, discriminators + losses, z2hidden: ONE launch)",
                           "bound": "hbm", "achieved": heads_bytes / (heads_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": heads_bytes / (heads_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "launch_ms": heads_ms,
                           "algorithmic_bytes": heads_bytes,
                           "note": "launch/latency bound as SURVEY 8d predicts; inside the kernel every CTA streams the full weight "
                                   "matrices from L2 (profiles/probes/heads_timeline.py)"},
                          {"kernel": "encoder forward = embedding + dropout + 2 layers x (2 input-projection GEMMs + 1 persistent "
                                     "bidirectional tcgen05 recurrence kernel)",
                           "bound": "latency", "achieved": enc_flops / (enc_ms * 1e-3) / 1e12, "peak": peaks["bf16_tflops"],
                           "unit": "TFLOP/s", "frac": enc_flops / (enc_ms * 1e-3) / 1e12 / peaks["bf16_tflops"], "launch_ms": enc_ms,
                           "algorithmic_flops": enc_flops, "sequential_steps": 2 * T,
                           "us_per_sequential_step": enc_ms * 1e3 / (2 * T)}]}
    line = {"metric": "train_tokens_per_sec", "value": tok_sum / (dev_ms * 1e-3), "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "seq_len": T, "parallelism": f"dp{world}",
                       "l2": "flushed (256 MiB write) between timed steps", "cuda_graph": not args.no_graph,
                       "tokens": "valid tokens (sum of lengths, incl. SOS/EOS)"},
            "padded_tokens_per_sec": B * world * T * args.steps / (dev_ms * 1e-3),
            "e2e": {"value": e2e_tok / e2e_s, "unit": "tokens/s", "h2d_bytes_per_step": eng.h2d_bytes_per_step,
                    "d2h_bytes_per_step": eng.d2h_bytes_per_step, "ms_per_step": e2e_s * 1e3 / args.steps},
            "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
            "roofline": roof, "clocks": clk, "loss_after_warmup": loss_warm, "loss_last": loss_last}
    if not args.no_cpu_baseline and world == 1 and args.workload == "cfg2":
        tps, sps, sample = cpu_port_steps(3, 1, budget_s=25.0)
        line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port", "sample": sample,
                                "s_per_step": sps}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
