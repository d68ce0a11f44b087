"""The slice of the reference's TrainConfig that the hot path consumes.

The reference's TrainConfig (/root/reference/modules/util/config/TrainConfig.py:265-995) has ~200
fields; the kernels' host code reads the ones below, by attribute name, so either object can be
passed.  Defaults are the reference's (`TrainConfig.default_values`, :759-995).
"""
from __future__ import annotations

from dataclasses import dataclass, field

from ..enum import EMAMode, LearningRateScaler, LossScaler, LossWeight, Optimizer, PeftType, TimestepDistribution


class _ModelTypeStub:
    """Stands in for ModelType where only the inpainting predicates are consulted
    (/root/reference/modules/util/enum/ModelType.py:92-96)."""

    def __init__(self, conditioning_image_input: bool = False, mask_input: bool = False):
        self._cond, self._mask = conditioning_image_input, mask_input

    def has_conditioning_image_input(self) -> bool:
        return self._cond

    def has_mask_input(self) -> bool:
        return self._mask


@dataclass
class TrainOptimizerConfig:
    """The fields of the reference's TrainOptimizerConfig (TrainConfig.py:34-262) that `create_optimizer` reads for
    ADAM / ADAMW (create.py:481-534), ADAFACTOR (:914-936) and CAME (:939-951); None means "use the optimizer's default",
    as there."""
    optimizer: object = Optimizer.ADAMW
    beta1: float | None = None
    beta2: float | None = None
    beta3: float | None = None
    eps: float | None = None
    eps2: float | None = None
    clip_threshold: float | None = None
    decay_rate: float | None = None
    scale_parameter: bool | None = None
    relative_step: bool | None = None
    warmup_init: bool | None = None
    weight_decay: float | None = None
    amsgrad: bool | None = False
    foreach: bool | None = False
    fused: bool | None = False
    maximize: bool | None = False
    capturable: bool | None = False
    differentiable: bool | None = False
    fused_back_pass: bool = False
    stochastic_rounding: bool = False


@dataclass
class TrainConfig:
    train_device: str = "cuda"
    batch_size: int = 1
    gradient_accumulation_steps: int = 1
    model_type: object = field(default_factory=_ModelTypeStub)
    # LoRA
    peft_type: object = PeftType.LORA
    lora_rank: int = 16
    lora_alpha: float = 1.0
    lora_decompose: bool = False
    lora_decompose_norm_epsilon: bool = True
    lora_decompose_output_axis: bool = False
    lora_layers: str = ""
    dropout_probability: float = 0.0
    # noise
    offset_noise_weight: float = 0.0
    perturbation_noise_weight: float = 0.0
    timestep_distribution: object = TimestepDistribution.UNIFORM
    min_noising_strength: float = 0.0
    max_noising_strength: float = 1.0
    noising_weight: float = 0.0
    noising_bias: float = 0.0
    timestep_shift: float = 1.0
    dynamic_timestep_shifting: bool = False
    # loss
    masked_training: bool = False
    unmasked_weight: float = 0.1
    normalize_masked_area_loss: bool = False
    masked_prior_preservation_weight: float = 0.0
    mse_strength: float = 1.0
    mae_strength: float = 0.0
    log_cosh_strength: float = 0.0
    vb_loss_strength: float = 1.0
    loss_weight_fn: object = LossWeight.CONSTANT
    loss_weight_strength: float = 5.0
    loss_scaler: object = LossScaler.NONE
    # optimizer / ema (consumed by the trainer glue, not by kernels)
    learning_rate: float = 3e-6
    learning_rate_scaler: object = LearningRateScaler.NONE
    optimizer: TrainOptimizerConfig = field(default_factory=TrainOptimizerConfig)
    ema: object = EMAMode.OFF
    clip_grad_norm: float | None = 1.0
    ema_decay: float = 0.999
    ema_update_step_interval: int = 5

    @staticmethod
    def default_values() -> "TrainConfig":
        return TrainConfig()
