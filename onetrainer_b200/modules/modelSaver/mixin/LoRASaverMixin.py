"""On-disk LoRA key formats — mirror of /root/reference/modules/modelSaver/mixin/LoRASaverMixin.py:38-106 (SURVEY.md §8f
rank 4) with DtypeModelSaverMixin._convert_state_dict_dtype (mixin/DtypeModelSaverMixin.py:22-35) underneath.  Host-only:
it consumes `LoRAModuleWrapper.state_dict()` (keys `{prefix}.{module path}.lora_down.weight | lora_up.weight | alpha |
dora_scale`) and writes safetensors files other tools can read.

Formats (`_save`, reference :88-106):
  INTERNAL             `<destination>/lora/lora.safetensors`, fp dtype untouched
  LEGACY_SAFETENSORS   "legacy diffusers" keys — and SAFETENSORS while `enable_omi_format` is off, as in the reference
  SAFETENSORS + OMI    OMI keys (`convert_to_omi`)
  DIFFUSERS            NotImplementedError, as in the reference (:97-98)

The key conversion itself is NOT in the reference's tree: it lives in `omi_model_standards.convert.lora.convert_lora_util`
(git+https://github.com/Open-Model-Initiative/OMI-Model-Standards.git@4ad235c, requirements-global.txt:29), which is not
installed here and cannot be fetched.  What is restated below is the published LEGACY naming rule that library applies to
keys that already carry diffusers module paths (the only kind this path produces): the adapter prefix, an underscore, the
module path with every '.' replaced by '_', then the unchanged leaf suffix —
    lora_unet.down_blocks.1.attentions.0.transformer_blocks.0.attn1.to_q.lora_down.weight
 -> lora_unet_down_blocks_1_attentions_0_transformer_blocks_0_attn1_to_q.lora_down.weight
(the kohya-ss convention every UI loader expects).  The OMI layout needs that library's per-model key tables (original
checkpoint names); asking for it raises NotImplementedError naming the dependency instead of guessing.
"""
from __future__ import annotations

import os
from abc import ABCMeta, abstractmethod
from dataclasses import dataclass
from enum import Enum
from pathlib import Path

import torch
from torch import Tensor

LEAF_SUFFIXES = (".lora_down.weight", ".lora_up.weight", ".alpha", ".dora_scale",
                 ".hada_w1_a", ".hada_w1_b", ".hada_w2_a", ".hada_w2_b")


class ModelFormat(Enum):
    """The members `_save` switches on (/root/reference/modules/util/enum/ModelFormat.py)."""
    DIFFUSERS = "DIFFUSERS"
    SAFETENSORS = "SAFETENSORS"
    LEGACY_SAFETENSORS = "LEGACY_SAFETENSORS"
    INTERNAL = "INTERNAL"


@dataclass(frozen=True)
class LoraConversionKeySet:
    """One adapter prefix of a model family (the slice of omi_model_standards' LoraConversionKeySet this path uses)."""
    diffusers_prefix: str                 # as LoRAModuleWrapper writes it: "lora_unet", "lora_te1", "lora_transformer"
    legacy_diffusers_prefix: str | None = None

    @property
    def legacy(self) -> str:
        return self.legacy_diffusers_prefix or self.diffusers_prefix


def sdxl_lora_key_sets() -> list[LoraConversionKeySet]:
    """The prefixes StableDiffusionXLLoRASetup creates (StableDiffusionXLLoRASetup.py:103-131)."""
    return [LoraConversionKeySet("lora_te1"), LoraConversionKeySet("lora_te2"), LoraConversionKeySet("lora_unet")]


def flux_lora_key_sets() -> list[LoraConversionKeySet]:
    return [LoraConversionKeySet("lora_te1"), LoraConversionKeySet("lora_te2"), LoraConversionKeySet("lora_transformer")]


def _split_leaf(key: str) -> tuple[str, str] | None:
    for suffix in LEAF_SUFFIXES:
        if key.endswith(suffix):
            return key[: -len(suffix)], suffix
    return None


def convert_to_legacy_diffusers(state_dict: dict[str, Tensor], key_sets: list[LoraConversionKeySet]) -> dict[str, Tensor]:
    """Internal keys -> legacy diffusers keys (see the module docstring).  Keys outside the key sets (bundled embeddings,
    `bundle_emb.*`) pass through unchanged."""
    out = {}
    for key, value in state_dict.items():
        split = _split_leaf(key)
        hit = next((ks for ks in key_sets if key.startswith(ks.diffusers_prefix + ".")), None) if split else None
        if hit is None:
            out[key] = value
            continue
        path, suffix = split
        module_path = path[len(hit.diffusers_prefix) + 1:]
        out[f"{hit.legacy}_{module_path.replace('.', '_')}{suffix}"] = value
    if len(out) != len(state_dict):
        raise ValueError("two LoRA keys collapse onto one legacy key")
    return out


def convert_legacy_diffusers_to_internal(state_dict: dict[str, Tensor], internal_keys) -> dict[str, Tensor]:
    """The inverse, for loading: '_' is ambiguous in a legacy key, so the mapping is built from the internal keys the
    target wrapper knows (`wrapper.state_dict().keys()`); unknown keys pass through (the wrapper parks them on dummies)."""
    table = {}
    for ik in internal_keys:
        split = _split_leaf(ik)
        if split:
            prefix, _, module_path = split[0].partition(".")
            table[f"{prefix}_{module_path.replace('.', '_')}{split[1]}"] = ik
    return {table.get(k, k): v for k, v in state_dict.items()}


def convert_to_omi(state_dict, key_sets):
    raise NotImplementedError(
        "the OMI LoRA key layout needs the per-model key tables of omi_model_standards (OMI-Model-Standards @ 4ad235c, "
        "requirements-global.txt:29), which is not installed; save with enable_omi_format=False (legacy diffusers keys)")


class DtypeModelSaverMixin:
    def _convert_state_dict_dtype(self, state_dict: dict, dtype: torch.dtype | None) -> dict:
        """Everything to the CPU in `dtype` (None keeps each tensor's own), nested dicts included (reference :22-35)."""
        return {k: (self._convert_state_dict_dtype(v, dtype) if isinstance(v, dict) else v.to(device="cpu", dtype=dtype))
                for k, v in state_dict.items()}

    def _create_safetensors_header(self, model, state_dict: dict[str, Tensor] | None = None) -> dict[str, str]:
        """The reference assembles ModelSpec + git + kohya fields (:63-96); this path has neither a ModelSpec nor a git
        checkout to describe, so only the fields loaders key on are written."""
        header = {"format": "pt"}
        if getattr(model, "sdxl", False):
            header["ss_base_model_version"] = "sdxl_"       # the Automatic1111 webui reads this to pick the model family
        return header


class LoRASaverMixin(DtypeModelSaverMixin, metaclass=ABCMeta):
    @abstractmethod
    def _get_convert_key_sets(self, model) -> list[LoraConversionKeySet] | None:
        pass

    @abstractmethod
    def _get_state_dict(self, model) -> dict[str, Tensor]:
        pass

    def __write(self, model, destination: str, dtype: torch.dtype | None, convert) -> None:
        from safetensors.torch import save_file
        save_state_dict = self._convert_state_dict_dtype(self._get_state_dict(model), dtype)
        key_sets = self._get_convert_key_sets(model)
        if key_sets is not None:
            save_state_dict = convert(save_state_dict, key_sets)
        save_state_dict = {k: v.contiguous() for k, v in save_state_dict.items()}
        os.makedirs(Path(destination).parent.absolute(), exist_ok=True)
        save_file(save_state_dict, destination, self._create_safetensors_header(model, save_state_dict))

    def _save(self, model, output_model_format, output_model_destination: str, dtype: torch.dtype | None,
              enable_omi_format: bool = False) -> None:
        fmt = output_model_format.name if isinstance(output_model_format, Enum) else str(output_model_format)
        if fmt == "DIFFUSERS":
            raise NotImplementedError
        if fmt == "SAFETENSORS":
            self.__write(model, output_model_destination, dtype, convert_to_omi if enable_omi_format else convert_to_legacy_diffusers)
        elif fmt == "LEGACY_SAFETENSORS":
            self.__write(model, output_model_destination, dtype, convert_to_legacy_diffusers)
        elif fmt == "INTERNAL":
            os.makedirs(output_model_destination, exist_ok=True)
            # reference :76-86 routes INTERNAL through the OMI writer; without that library the internal backup keeps the
            # wrapper's own keys, which is what load_state_dict consumes
            self.__write(model, os.path.join(output_model_destination, "lora", "lora.safetensors"), None, lambda sd, ks: sd)
        else:
            raise ValueError(f"unknown model format {fmt}")


class StandInLoRASaver(LoRASaverMixin):
    """`StableDiffusionXLLoRASaver` (modelSaver/stableDiffusionXL/StableDiffusionXLLoRASaver.py:12-59) for the fields the
    harness models carry: the UNet / transformer adapter and passthrough keys."""

    def __init__(self, key_sets: list[LoraConversionKeySet] | None = None):
        self._key_sets = key_sets if key_sets is not None else sdxl_lora_key_sets() + [LoraConversionKeySet("lora_transformer")]

    def _get_convert_key_sets(self, model):
        return self._key_sets

    def _get_state_dict(self, model) -> dict[str, Tensor]:
        state_dict = {}
        for name in ("text_encoder_1_lora", "text_encoder_2_lora", "unet_lora", "transformer_lora"):
            wrapper = getattr(model, name, None)
            if wrapper is not None:
                state_dict |= wrapper.state_dict()
        if getattr(model, "lora_state_dict", None) is not None:
            state_dict |= model.lora_state_dict
        return state_dict

    def save(self, model, output_model_format, output_model_destination: str, dtype: torch.dtype | None) -> None:
        self._save(model, output_model_format, output_model_destination, dtype)
