"""`.npy` segment dataset feeding scripts/resynthesise_dataset.py — mirror of the reference's
data/general.py:9-57 (on-disk layout: <root>/<split>/{audio,control}/{audio,control}_<name>.npy plus
<root>/data_mean.npy and data_std.npy, each [n_features, 1])."""
import os

import numpy as np
import torch


class GeneralDataset(torch.utils.data.Dataset):
    def __init__(self, path: str, split: str = "train", load_to_memory: bool = True):
        super().__init__()
        self.load_to_memory = load_to_memory
        self.split_path = os.path.join(path, split)
        self.data_list = sorted(f[len("audio_"):] for f in os.listdir(os.path.join(self.split_path, "audio"))
                                if f.endswith(".npy"))
        self.data_mean = np.load(os.path.join(path, "data_mean.npy"))
        self.data_std = np.load(os.path.join(path, "data_std.npy"))
        if load_to_memory:
            self.audio = [self._load("audio", n) for n in self.data_list]
            self.control = [self._load("control", n) for n in self.data_list]

    def _load(self, kind, name):
        return np.load(os.path.join(self.split_path, kind, "%s_%s" % (kind, name)))

    def __len__(self):
        return len(self.data_list)

    def __getitem__(self, idx):
        name = self.data_list[idx]
        audio = self.audio[idx] if self.load_to_memory else self._load("audio", name)
        control = self.control[idx] if self.load_to_memory else self._load("control", name)
        denorm = control * self.data_std + self.data_mean
        return {"audio": audio, "f0": denorm[0:1, :], "amp": denorm[1:2, :], "control": control,
                "name": os.path.splitext(os.path.basename(name))[0]}
