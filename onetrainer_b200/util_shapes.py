"""Shape helpers shared by the host mirrors."""
from __future__ import annotations


def flatten_chw(shape) -> tuple[int, int, int]:
    """(B, C, H*W...) of a [B, C, *spatial] tensor; [B, L] and [B] are treated as C=1."""
    shape = tuple(int(s) for s in shape)
    if len(shape) == 0:
        raise ValueError("scalar tensors have no batch dimension")
    B = shape[0]
    if len(shape) == 1:
        return B, 1, 1
    if len(shape) == 2:
        return B, 1, shape[1]
    hw = 1
    for s in shape[2:]:
        hw *= s
    return B, shape[1], hw
