"""Reader for tests/golden/*.npz (written by oracle/make_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.meta = json.loads(str(z["__meta__"]))
        bf = set(json.loads(str(z["__bf16__"]))) if "__bf16__" in z.files else set()
        self._t = {}
        for k in z.files:
            if k.startswith("__"):
                continue
            a = z[k]
            if k in bf:
                self._t[k] = torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)
            else:
                self._t[k] = torch.from_numpy(a.copy())

    def __getitem__(self, k):
        return self._t[k]

    def __contains__(self, k):
        return k in self._t

    @property
    def state_dict(self):
        return {k[3:]: v for k, v in self._t.items() if k.startswith("sd.")}

    @property
    def dtype(self):
        return getattr(torch, self.meta["dtype"])
