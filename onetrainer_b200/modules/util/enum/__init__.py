"""Enumerations the hot path reads from a TrainConfig.

Values are the same strings the reference's enums carry
(/root/reference/modules/util/enum/{TimestepDistribution,LossWeight,LossScaler}.py), so a config
object from either code base works: the kernels' host code compares by `enum_name()`.
"""
from enum import Enum


class _StrEnum(Enum):
    def __str__(self) -> str:
        return self.value


TimestepDistribution = _StrEnum("TimestepDistribution", {k: k for k in ("UNIFORM", "SIGMOID", "LOGIT_NORMAL", "HEAVY_TAIL", "COS_MAP")})
LossWeight = _StrEnum("LossWeight", {k: k for k in ("CONSTANT", "P2", "MIN_SNR_GAMMA", "DEBIASED_ESTIMATION", "SIGMA")})
LossScaler = _StrEnum("LossScaler", {k: k for k in ("NONE", "BATCH", "GRADIENT_ACCUMULATION", "BOTH")})
PeftType = _StrEnum("PeftType", {k: k for k in ("LORA", "LOHA")})
LearningRateScaler = _StrEnum("LearningRateScaler", {k: k for k in ("NONE", "BATCH", "GRADIENT_ACCUMULATION", "BOTH")})
EMAMode = _StrEnum("EMAMode", {k: k for k in ("OFF", "GPU", "CPU")})
Optimizer = _StrEnum("Optimizer", {k: k for k in ("ADAM", "ADAMW", "ADAFACTOR", "CAME")})     # the ones with a fused step


def enum_name(x) -> str:
    """Name of an enum member from this package, from the reference, or a plain string."""
    if isinstance(x, Enum):
        return x.name
    return str(x)
