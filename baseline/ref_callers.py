"""The reference's OWN family callers — `BaseStableDiffusionSetup.predict / calculate_loss`
(modules/modelSetup/BaseStableDiffusionSetup.py:135-325), `BaseStableDiffusionXLSetup.predict / calculate_loss`
(BaseStableDiffusionXLSetup.py:179-373) and `BaseFluxSetup.predict / calculate_loss` (BaseFluxSetup.py:193-390) — run
UNMODIFIED from baseline/_ref, in two arms:

* reference arm: the files as they are, on the reference's own mixins and LoRAModuleWrapper;
* drop-in arm:  the SAME source files executed a second time with the four seam modules
  (`modules.modelSetup.mixin.ModelSetup{Noise,Diffusion,FlowMatching,DiffusionLoss}Mixin`) resolved to this repo's
  mirrors — INTEGRATION.md §2(b)'s `sys.modules` redirection — so the unmodified `predict()` bodies call
  `self._create_noise`, `self._get_timestep_discrete`, `self._add_noise_discrete`, `self._diffusion_losses` /
  `self._flow_matching_losses` of onetrainer_b200, and the network they call is hooked by this repo's LoRAModuleWrapper.

BENCH / TEST INFRASTRUCTURE ONLY (nothing under onetrainer_b200/ imports this file).  The model objects are the
reference's real `StableDiffusionModel` / `StableDiffusionXLModel` / `FluxModel` classes with their network, VAE config
and scheduler fields filled by the seeded stand-ins (the wrapped diffusers networks are outside the path and not
installed); text-encoder outputs come cached in the batch, as with the reference's latent/text caching.

Import shims (on top of baseline/ref_arm.py's two): `diffusers` is absent, so a stub package answers every
`from diffusers... import Name` with an empty placeholder class (the callers use those names for type annotations and
`isinstance` checks in code this path never reaches); transformers 5 dropped `tokenization_utils.Trie`
(ModelSetupEmbeddingMixin.py:18 imports it for textual-inversion code that is not on the path).
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys
import types

from . import ref_arm

SEAM_MODULES = ("modelSetup.mixin.ModelSetupNoiseMixin", "modelSetup.mixin.ModelSetupDiffusionMixin",
                "modelSetup.mixin.ModelSetupFlowMatchingMixin", "modelSetup.mixin.ModelSetupDiffusionLossMixin")
CALLER_MODULES = ("modules.modelSetup.BaseStableDiffusionSetup", "modules.modelSetup.BaseStableDiffusionXLSetup",
                  "modules.modelSetup.BaseFluxSetup")


class _Placeholders(types.ModuleType):
    """A module whose every attribute is an empty class named after it."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        t = type(name, (), {})
        setattr(self, name, t)
        return t


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, top: str):
        self.top = top

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] == self.top:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Placeholders(spec.name)
        m.__path__ = []
        m._otb_stub = True
        return m

    def exec_module(self, module):
        pass


def install_caller_shims() -> None:
    ref_arm.install_shims()
    try:
        import diffusers
        real = not getattr(diffusers, "_otb_stub", False) and hasattr(diffusers, "__file__") and diffusers.__file__
    except Exception:  # noqa: BLE001
        real = False
    if not real and not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        for k in [k for k in sys.modules if k.split(".")[0] == "diffusers"]:
            del sys.modules[k]
        sys.meta_path.insert(0, _StubFinder("diffusers"))
    # transformers re-creates its deprecated alias modules lazily: touch everything the reference imports first
    from transformers import CLIPTextModel, CLIPTextModelWithProjection, CLIPTokenizer, PreTrainedTokenizer, T5EncoderModel, T5Tokenizer  # noqa: F401
    tu = sys.modules.get("transformers.tokenization_utils")
    if tu is None or not hasattr(tu, "Trie"):
        stub = types.ModuleType("transformers.tokenization_utils")
        stub.PreTrainedTokenizer = PreTrainedTokenizer
        stub.Trie = type("Trie", (), {})
        sys.modules["transformers.tokenization_utils"] = stub


def available() -> bool:
    return ref_arm.available() and (ref_arm.REF_DIR / "modules" / "modelSetup" / "BaseStableDiffusionXLSetup.py").exists()


@contextlib.contextmanager
def _seams_redirected():
    """INTEGRATION.md §2(b) for the duration of one import: `modules.<seam>` resolves to onetrainer_b200's mirror."""
    saved = {}
    try:
        for name in SEAM_MODULES:
            key = "modules." + name
            saved[key] = sys.modules.get(key)
            sys.modules[key] = importlib.import_module("onetrainer_b200.modules." + name)
        yield
    finally:
        for key, mod in saved.items():
            if mod is None:
                sys.modules.pop(key, None)
            else:
                sys.modules[key] = mod


_cache: dict = {}


def load(dropin: bool):
    """Namespace with the three caller classes (made concrete: the abstract set-up methods that are not on the path are
    no-ops) and the reference's model classes / enums.  `dropin=True`: the same source files, executed with the seam
    modules redirected to this repo."""
    if dropin in _cache:
        return _cache[dropin]
    if not available():
        raise RuntimeError("baseline/_ref does not hold the family callers: run tools/make_baseline_ref.py in the build container")
    install_caller_shims()
    R = ref_arm.load()
    ns = types.SimpleNamespace(R=R, dropin=dropin)
    mods = {}
    if not dropin:
        for name in CALLER_MODULES:
            mods[name] = importlib.import_module(name)
    else:
        for name in CALLER_MODULES:
            importlib.import_module(name)          # the reference's own copy first: all shared (non-seam) imports are cached
        with _seams_redirected():
            for name in CALLER_MODULES:
                path = ref_arm.REF_DIR / (name.replace(".", "/") + ".py")
                spec = importlib.util.spec_from_file_location("otb_dropin." + name, path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)       # the unmodified file; its `from modules...Mixin import` lines now find ours
                mods[name] = mod
    from modules.model.FluxModel import FluxModel
    from modules.model.StableDiffusionModel import StableDiffusionModel
    from modules.model.StableDiffusionXLModel import StableDiffusionXLModel
    from modules.util.enum.DataType import DataType
    from modules.util.enum.ModelType import ModelType
    from modules.util.TrainProgress import TrainProgress

    def concrete(cls):
        noop = lambda self, *a, **k: None     # noqa: E731
        return type("Concrete" + cls.__name__, (cls,), {m: noop for m in cls.__abstractmethods__})

    ns.SDSetup = concrete(mods[CALLER_MODULES[0]].BaseStableDiffusionSetup)
    ns.SDXLSetup = concrete(mods[CALLER_MODULES[1]].BaseStableDiffusionXLSetup)
    ns.FluxSetup = concrete(mods[CALLER_MODULES[2]].BaseFluxSetup)
    ns.caller_files = {k: m.__file__ for k, m in mods.items()}
    ns.mixin_owner = {c.__name__: c.__module__ for c in ns.SDXLSetup.__mro__ if c.__name__.startswith("ModelSetup")}
    ns.StableDiffusionModel, ns.StableDiffusionXLModel, ns.FluxModel = StableDiffusionModel, StableDiffusionXLModel, FluxModel
    ns.DataType, ns.ModelType, ns.TrainProgress = DataType, ModelType, TrainProgress
    _cache[dropin] = ns
    return ns


def caller_config(ns, kind: str, **overrides):
    """The reference's TrainConfig defaults with the fields the LoRA presets set for this path
    (training_presets/#sdxl 1.0 LoRA.json, #sd 1.5 LoRA.json, #flux LoRA.json): text encoders frozen (their outputs come
    cached in the batch), the family's model type."""
    cfg = ref_arm.reference_config(**overrides)
    cfg.model_type = ns.ModelType.FLUX_DEV_1 if "flux" in kind else ns.ModelType.STABLE_DIFFUSION_XL_10_BASE if kind in ("sdxl", "tiny") \
        else ns.ModelType.STABLE_DIFFUSION_15
    cfg.text_encoder.train = False
    cfg.text_encoder_2.train = False
    cfg.text_encoder.train_embedding = False
    cfg.text_encoder_2.train_embedding = False
    return cfg


class _Config(dict):
    """diffusers' FrozenDict: item and attribute access (the callers use both: `config['scaling_factor']`,
    `config.prediction_type`, `config.guidance_embeds`)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class _Holder:
    def __init__(self, **config):
        self.config = _Config(config)


def make_model(ns, kind: str, network, scheduler, train_dtype, device, lora_wrapper=None):
    """The reference's real model class for `kind` ('sd15' | 'sdxl' | 'flux' | tiny variants), fields filled from the
    stand-ins: what StableDiffusion*ModelLoader + *LoRASetup.setup_model / setup_optimizations leave behind."""
    import torch
    flux = "flux" in kind
    sdxl = kind in ("sdxl", "tiny")
    if flux:
        model = ns.FluxModel(ns.ModelType.FLUX_DEV_1)
        model.transformer = network
        model.vae = _Holder(scaling_factor=0.3611, shift_factor=0.1159)
        model.transformer_lora = lora_wrapper
        if not isinstance(network.config, _Config):
            network.config = _Config(network.config)
    elif sdxl:
        model = ns.StableDiffusionXLModel(ns.ModelType.STABLE_DIFFUSION_XL_10_BASE)
        model.unet = network
        model.vae = _Holder(scaling_factor=0.13025)
        model.unet_lora = lora_wrapper
    else:
        model = ns.StableDiffusionModel(ns.ModelType.STABLE_DIFFUSION_15)
        model.unet = network
        model.vae = _Holder(scaling_factor=0.18215)
        model.unet_lora = lora_wrapper
    if not isinstance(scheduler.config, _Config):
        scheduler.config = _Config(scheduler.config)
    model.noise_scheduler = scheduler
    dt = {torch.float32: ns.DataType.FLOAT_32, torch.bfloat16: ns.DataType.BFLOAT_16}[train_dtype]
    model.train_dtype = dt
    dev = torch.device(device)
    model.autocast_context = torch.autocast(device_type="cuda", dtype=train_dtype) if dev.type == "cuda" and train_dtype != torch.float32 \
        else contextlib.nullcontext()
    model.train_progress = ns.TrainProgress()
    return model


def caller_batch(kind: str, batch: dict) -> dict:
    """The harness' synthetic batch under the key names the reference's data loaders emit
    (modules/dataLoader/StableDiffusionXLBaseDataLoader.py:174-193 etc.): cached text-encoder states per encoder."""
    import torch
    b = dict(batch)
    B = batch["latent_image"].shape[0]
    dev = batch["latent_image"].device
    tok = torch.zeros((B, 77), dtype=torch.long, device=dev)
    if "flux" in kind:
        b.setdefault("tokens_1", tok)
        b.setdefault("tokens_2", tok)
    elif kind in ("sdxl", "tiny"):
        h = batch["text_encoder_hidden_state"]
        split = h.shape[-1] * 3 // 8                    # 768 of 2048 (CLIP-L | OpenCLIP-bigG), same ratio on the tiny stand-in
        b["text_encoder_1_hidden_state"], b["text_encoder_2_hidden_state"] = h[..., :split], h[..., split:]
        b["tokens_1"] = b["tokens_2"] = tok
    else:
        b["tokens"] = tok
    return b
