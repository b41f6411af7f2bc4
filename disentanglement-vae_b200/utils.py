"""Config / seeding helpers with the reference's semantics (vae/utils.py:13-88)."""
import json
import random

import numpy as np
import torch

# schema of the params JSON (vae/utils.py:49-77); `bow_encoder` and `mi_loss` are required by the
# reference's validator although `config_example.json` omits them.
VALID_PARAMS = {
    "name": str, "random_seed": int, "data_dir": str, "combined_dataset": bool, "dataset_minibatch_ratios": dict,
    "checkpoint_dir": str, "glove_path": str, "num_train_examples": int, "lowercase": bool, "reverse_input": bool,
    "embedding_dim": int, "hidden_dim": int, "num_rnn_layers": int, "bidirectional_encoder": bool,
    "bow_encoder": bool, "latent_dims": dict, "epochs": int, "batch_size": int, "learn_rate": float,
    "encoder_dropout": float, "decoder_dropout": float, "teacher_forcing_prob": float, "lambdas": dict,
    "adversarial_loss": bool, "mi_loss": bool, "train": bool, "validate": bool, "test": bool}

# keys this package adds; the reference's validator only warns about unknown keys (utils.py:85-87)
OPTIONAL_PARAMS = {"backend": str, "world_size": int}


def set_seed(seed):
    """vae/utils.py:13-19."""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def validate_params(params):
    """vae/utils.py:48-88: same required keys, same types, same ValueError on violation."""
    for key, typ in VALID_PARAMS.items():
        if key not in params:
            raise ValueError(f"parameter file missing '{key}'")
        if not isinstance(params[key], typ):
            raise ValueError(f"Parameter '{key}' of incorrect type!")
    for key in params:
        if key not in VALID_PARAMS and key not in OPTIONAL_PARAMS:
            print(f"WARNING: Ignoring unused parameter '{key}' in parameter file.")
    if "total" not in params["latent_dims"]:
        raise ValueError("latent_dims needs a 'total' entry")
    if "default" not in params["lambdas"]:
        raise ValueError("lambdas needs a 'default' entry")
    for k, v in params["lambdas"].items():
        if not (isinstance(v, (int, float)) or v == "cyclic"):
            raise ValueError(f"lambda '{k}' must be a number or \"cyclic\"")


def load_params(path):
    with open(path) as f:
        params = json.load(f)
    validate_params(params)
    return params


def kl_weights_for_step(params, step, total_steps):
    """run.py:230-236: resolve each lambda, evaluating "cyclic" at this global step."""
    from .losses import get_cyclic_kl_weight
    return {k: (get_cyclic_kl_weight(step, total_steps) if v == "cyclic" else v)
            for k, v in params["lambdas"].items()}


# ---- on-disk formats (run.py:166-194, 624-630; vae/utils.py:147-175) ------------------------------------------------
def load_latest_checkpoint(model, optimizer, checkpoint_dir, map_location=None):
    """vae/utils.py:147-175: load the highest-epoch `model_{epoch}.pt` of `checkpoint_dir` into model and optimizer
    (a torch optimizer or an `engine.TrainEngine`, which speaks the same state-dict format); returns
    (model, optimizer, next_epoch, file name) or (model, optimizer, 0, None) when there is no checkpoint."""
    import os
    ckpts = [f for f in os.listdir(checkpoint_dir) if f.endswith(".pt")]
    if not ckpts:
        return model, optimizer, 0, None
    # the reference starts its search at epoch 0 with a strict '>', so with several files it keeps ckpts[0] unless a later
    # epoch is larger; same rule here
    latest, latest_epoch = ckpts[0], 0
    for f in ckpts:
        epoch = int(f.replace("model_", "").replace(".pt", ""))
        if epoch > latest_epoch:
            latest, latest_epoch = f, epoch
    ckpt = torch.load(os.path.join(checkpoint_dir, latest), map_location=map_location)
    model.load_state_dict(ckpt["model_state_dict"])
    optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return model, optimizer, ckpt["epoch"] + 1, latest


def log_params(params_dict, example_ids, logdir, dataset_name, epoch):
    """run.py:166-194: metadata/ordered_ids/{dataset}_{epoch}.log (one id per line) and
    metadata/{z,mu,logvar}/{dataset}_{latent}_{epoch}.log (CSV, one row per example, "%.4f")."""
    import csv
    import os
    metadata_dir = os.path.join(logdir, "metadata")
    ids_dir = os.path.join(metadata_dir, "ordered_ids")
    os.makedirs(ids_dir, exist_ok=True)
    with open(os.path.join(ids_dir, f"{dataset_name}_{epoch}.log"), "w") as f:
        for i in example_ids:
            f.write(f"{i}\n")
    for latent_name, per_param in params_dict.items():
        for param_name, values in per_param.items():
            param_dir = os.path.join(metadata_dir, param_name)
            os.makedirs(param_dir, exist_ok=True)
            with open(os.path.join(param_dir, f"{dataset_name}_{latent_name}_{epoch}.log"), "w") as f:
                writer = csv.writer(f, delimiter=",")
                for value in values:
                    writer.writerow([f"{dim:.4f}" for dim in value])


class LatentLog:
    """Device-side accumulation of the per-example latents the reference dumps every epoch (run.py:279-283: a
    `.detach().cpu().tolist()` per tensor per step, i.e. 3 * n_spaces device syncs per step).  `append(plan, ids)` enqueues
    ONE asynchronous copy of (z, mu, logvar) [3,B,Z] into a pinned host chunk -- no sync on the training stream -- and
    `flush()` waits once, then writes the reference's files through `log_params`."""

    def __init__(self, space_names, space_dims, chunk_steps=256):
        self.names, self.dims = list(space_names), list(space_dims)
        self.Z = sum(self.dims)
        self.chunk_steps = chunk_steps
        self._chunks, self._ids, self._n = [], [], 0
        self._ev = None

    def append(self, plan, example_ids):
        B = plan.z.size(0)
        if not self._chunks or self._n == self.chunk_steps:
            self._chunks.append(torch.empty(self.chunk_steps, 3, B, self.Z, dtype=torch.float32).pin_memory())
            self._n = 0
        dst = self._chunks[-1][self._n]
        dst[0].copy_(plan.z, non_blocking=True)
        dst[1].copy_(plan.mu, non_blocking=True)
        dst[2].copy_(plan.logvar, non_blocking=True)
        self._n += 1
        self._ids.extend(example_ids)
        self._ev = torch.cuda.Event()
        self._ev.record()

    def collect(self):
        """{latent: {"z" | "mu" | "logvar": [N, zs] array}} in append order (waits for the pending copies)."""
        if self._ev is not None:
            self._ev.synchronize()
        parts = [c if i + 1 < len(self._chunks) else c[:self._n] for i, c in enumerate(self._chunks)]
        out = {n: {} for n in self.names}
        if not parts:
            return out
        allv = torch.cat(parts, 0)                       # [steps, 3, B, Z]
        off = 0
        for n, zs in zip(self.names, self.dims):
            for j, pname in enumerate(("z", "mu", "logvar")):
                out[n][pname] = allv[:, j, :, off:off + zs].reshape(-1, zs).numpy()
            off += zs
        return out

    def flush(self, logdir, dataset_name, epoch):
        log_params(self.collect(), self._ids, logdir, dataset_name, epoch)
        self._chunks, self._ids, self._n, self._ev = [], [], 0, None
