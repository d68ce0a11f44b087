"""Mapping from a torch.Generator to the kernels' Philox key.

The reference hands every mixin call the same per-step generator (seeded with `global_step`,
/root/reference/modules/modelSetup/BaseStableDiffusionXLSetup.py:189-191).  The kernels draw from
their own counter-based stream instead of torch's: key = the generator's
seed, offset = the generator's Philox offset, which is advanced on every draw so repeated calls
with one generator differ while a re-seeded generator reproduces — the semantics callers rely on.
"""
from __future__ import annotations

import torch


class PhiloxKey:
    """A stand-in for `torch.Generator` carrying just what the kernels consume (seed, offset, device).
    torch refuses to construct a CUDA generator while a CUDA graph is being captured; a captured step
    uses this instead — the draws are identical because only (seed, offset) ever reach the kernels."""

    def __init__(self, seed: int, device):
        self._seed, self._offset, self.device = int(seed), 0, torch.device(device)

    def manual_seed(self, seed: int):
        self._seed, self._offset = int(seed), 0
        return self

    def initial_seed(self) -> int:
        return self._seed

    def get_offset(self) -> int:
        return self._offset

    def set_offset(self, offset: int) -> None:
        self._offset = int(offset)


def draw_key(generator: torch.Generator | None, device=None) -> tuple[int, int]:
    """Returns (seed, offset) and consumes one tick of the generator."""
    if generator is None:
        generator = torch.cuda.default_generators[torch.device(device or "cuda").index or 0]
    seed = generator.initial_seed()
    try:
        offset = generator.get_offset()
        generator.set_offset(offset + 4)        # CUDA generators: offsets move in multiples of 4
    except RuntimeError:
        # CPU generators have no Philox offset: burn one draw so successive calls still differ
        offset = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=generator).item()) * 4
    return int(seed) & (2 ** 64 - 1), int(offset)
