from .schedules import *  # noqa: F401,F403
from .schedules import (SUPPORTED_NOISE_SCHEDULES, SUPPORTED_DISCRETIZATIONS,
                        SUPPORTED_SAMPLING_STEP_SCHEDULE, cosine_beta_schedule, linear_beta_schedule)
from .embeddings import (PositionalEmbedding, UntrainablePositionalEmbedding, SinusoidalEmbedding,
                         FourierEmbedding, UntrainableFourierEmbedding, SUPPORTED_TIMESTEP_EMBEDDING)
from .blocks import at_least_ndim, to_tensor, count_parameters, Mlp, GroupNorm1d


def set_seed(seed: int):
    """Seed python / numpy / torch (CPU + every CUDA device)."""
    import os, random
    import numpy as np
    import torch
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
