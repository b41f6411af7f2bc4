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
