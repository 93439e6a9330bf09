"""use_conv=True policy features (factor_net_ppo.py:108-130): cosine similarity of history slot 0 with slots
1..order_dim-1.  `cosine_features` is the autograd/torch form used on the PPO-update side; the sampling path uses
the CUDA reduction kernel (csrc/features.cu) through `cosine_features_cuda`."""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib


def cosine_features(epsilon: torch.Tensor, order_dim: int) -> torch.Tensor:
    B = epsilon.shape[0]
    flat = epsilon.reshape(B, order_dim, -1)
    first = flat[:, 0, :]
    return torch.cat([torch.nn.functional.cosine_similarity(flat[:, i, :], first, dim=-1).unsqueeze(-1)
                      for i in range(1, order_dim)], dim=-1)


def cosine_features_cuda(e0: torch.Tensor, cond: Optional[torch.Tensor], guidance: float, older: List[torch.Tensor],
                         order_dim: int, out: torch.Tensor, workspace: torch.Tensor, stream=None) -> torch.Tensor:
    """feat [B, order_dim-1] for the newest output `e0` (or the CFG pair e0/cond) against `older` (newest first)."""
    lib = _lib.load()
    B = e0.shape[0]
    N = e0.numel() // B
    if stream is None:
        stream = torch.cuda.current_stream(e0.device).cuda_stream
    rc = lib.consolver_cosine_features(_lib.dtype_code(e0.dtype), e0.data_ptr(),
                                       cond.data_ptr() if cond is not None else None, float(guidance),
                                       _lib.ptr_array([h.data_ptr() for h in older]), len(older) + 1, order_dim, B, N,
                                       workspace.data_ptr(), out.data_ptr(), stream)
    _lib.check(rc, "consolver_cosine_features")
    return out


def workspace_bytes(B: int, order_dim: int) -> int:
    return int(_lib.load().consolver_cosine_features_workspace(B, order_dim))
