"""B200-native disentangled sentence VAE: drop-in for `vae.model.build_vae` / `vae.losses` of
jvasilakes/disentanglement-vae, backed by hand-written sm_100a CUDA kernels."""
from . import _lib, losses, utils, model  # noqa: F401
from ._lib import DvaeError, launch_count  # noqa: F401
from .model import build_vae, VariationalSeq2Seq, VariationalEncoder, VariationalDecoder, Discriminator, Params  # noqa: F401
from .utils import set_seed, validate_params  # noqa: F401

__all__ = ["build_vae", "losses", "utils", "model", "set_seed", "validate_params", "DvaeError"]
