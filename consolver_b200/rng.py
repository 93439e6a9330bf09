"""Fused RNG: the sample kernel can generate the Exp(1) draw of `torch.multinomial` itself — bit-identical to
`torch.empty(B*A, K, device='cuda').exponential_(1)` on the default generator — instead of reading a q buffer
filled by a separate torch launch.  The host side only bookkeeps the generator: read (seed, offset), advance the
offset by what the torch launch would have consumed.  Seeds therefore keep reproducing the reference's actions.

A one-time self-check per device compares the kernel's draw with torch's; if the installed torch ever changes its
exponential_ kernel the check fails and the schedulers silently keep using the torch launch (still on the GPU,
still bit-exact — just one more kernel per step)."""
from __future__ import annotations

import ctypes as C
import os
import warnings
from typing import Dict, Tuple

import torch

from . import _lib

_verified: Dict[int, bool] = {}


_gens: Dict[int, torch.Generator] = {}


def generator(device: torch.device) -> torch.Generator:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    g = _gens.get(idx)
    if g is None:
        g = _gens[idx] = torch.cuda.default_generators[idx]
    return g


def take(device: torch.device, increment: int) -> Tuple[int, int]:
    """(seed, offset) for the next draw; advances the default generator like the torch launch would."""
    g = generator(device)
    seed, off = g.initial_seed(), g.get_offset()
    g.set_offset(off + increment)
    return seed, off


def draw_reference_check(device: torch.device, numel: int = 5 * 3 * 11) -> bool:
    """kernel draw == torch draw for `numel` elements (restores the generator state afterwards)"""
    lib = _lib.load()
    g = generator(device)
    saved = g.get_state()
    try:
        nthreads, inc = _lib.philox_plan(numel)
        seed, off = g.initial_seed(), g.get_offset()
        ref = torch.empty(numel, device=device, dtype=torch.float32).exponential_(1)
        consumed = g.get_offset() - off
        K = 11
        B = numel // K
        assert B * K == numel
        table = torch.full((1, K), 1.0 / K, device=device)
        av = torch.zeros(1, K, device=device)
        q_out = torch.empty(numel, device=device)
        outs = [torch.empty(B, device=device, dtype=torch.int64)] + [torch.empty(B, device=device) for _ in range(4)]
        coef = torch.empty(B, 4, device=device)
        rng = _lib.Rng(seed, off, None, nthreads)
        rc = lib.consolver_policy_sample_f32(table.data_ptr(), av.data_ptr(), None, None, C.byref(rng),
                                             q_out.data_ptr(), B, 1, K, 2, 0, 1, 0, *[o.data_ptr() for o in outs],
                                             coef.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "consolver_policy_sample_f32")
        return bool(torch.equal(ref, q_out)) and consumed == inc
    finally:
        g.set_state(saved)


def fused_rng_available(device: torch.device) -> bool:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    v = _verified.get(idx)
    if v is not None:
        return v
    if os.environ.get("CONSOLVER_FUSED_RNG", "1") != "1":
        _verified[idx] = False
        return False
    if idx not in _verified:
        if torch.cuda.is_current_stream_capturing():
            return False          # decide outside of a capture
        try:
            ok = draw_reference_check(device) and draw_reference_check(device, 33 * 11 * 4096 // 11 * 11)
        except Exception as e:  # noqa: BLE001
            warnings.warn(f"consolver_b200: fused RNG self-check raised {e!r}; using the torch exponential_ launch")
            ok = False
        if not ok:
            warnings.warn("consolver_b200: the in-kernel Exp(1) draw does not reproduce this torch build's "
                          "exponential_; falling back to the torch launch (results unchanged, one more kernel/step)")
        _verified[idx] = ok
    return _verified[idx]
