"""Per-step values that must change between replays of a captured CUDA graph.

A CUDA graph bakes kernel arguments in at capture time.  The hot path has four values that change
every step — the noise Philox key, the stochastic-rounding key, the EMA decay and the EMA on/off
flag — so, while a `DeviceStepState` is active, the host code writes them into a working copy, uploads
it through a ring of pinned slots right before each replay, and the kernels read them from device
memory (`rng_dev` / `dyn` in include/otb200.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N

_active: "DeviceStepState | None" = None


class DeviceStepState:
    NOISE = 0        # int64 words 0,1: noise seed, offset
    DYN = 2          # words 2..4: otb_optim_dyn {seed, offset, (float omd | int32 do_ema << 32)}

    def __init__(self, device):
        from .ops import PinnedRing
        self._np = np.zeros(8, dtype=np.int64)          # the host's working copy
        self._hosts = [torch.zeros(8, dtype=torch.int64).pin_memory() for _ in range(PinnedRing.SLOTS)]
        self._ring = PinnedRing()
        self.dev = torch.zeros(8, dtype=torch.int64, device=device)
        assert C.sizeof(N.OptimDyn) == 24

    def set_noise_key(self, seed: int, offset: int) -> None:
        self._np[0] = np.int64(np.uint64(seed & (2 ** 64 - 1)).astype(np.int64))
        self._np[1] = np.int64(offset)

    def set_optim(self, seed: int, offset: int, ema_one_minus_decay: float, do_ema: bool) -> None:
        self._np[2] = np.uint64(seed & (2 ** 64 - 1)).astype(np.int64)
        self._np[3] = np.int64(offset)
        packed = np.array([ema_one_minus_decay], dtype=np.float32).view(np.uint32)[0].astype(np.uint64) | (np.uint64(1 if do_ema else 0) << np.uint64(32))
        self._np[4] = packed.astype(np.int64)

    def upload(self) -> None:
        """Working copy -> device on the current stream, through a ring of pinned slots (the host may already be
        preparing the next step when the GPU executes the copy).  Issued BEFORE the graph replay, not captured."""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("DeviceStepState.upload() must run before graph.replay(), not inside the capture")
        slot = self._ring.acquire()
        self._hosts[slot].numpy()[:] = self._np
        self.dev.copy_(self._hosts[slot], non_blocking=True)
        self._ring.release(slot)

    @property
    def noise_ptr(self) -> torch.Tensor:
        return self.dev[self.NOISE:self.NOISE + 2]

    @property
    def dyn_ptr(self) -> torch.Tensor:
        return self.dev[self.DYN:self.DYN + 3]


def activate(state: DeviceStepState | None) -> None:
    global _active
    _active = state


def active() -> DeviceStepState | None:
    return _active
