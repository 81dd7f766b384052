"""Helpers the resynthesis script imports (reference utils/utils.py:20-23, utils/seed_all.py)."""
import os
import random

import numpy as np
import torch


def make_dir_if_not_exists(path):
    os.makedirs(path, exist_ok=True)


def seed_all(seed):
    np.random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
