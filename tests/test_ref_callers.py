"""the reference's OWN family callers, unmodified, on both sides of the seam.

`BaseStableDiffusionSetup.predict` (BaseStableDiffusionSetup.py:135-310), `BaseStableDiffusionXLSetup.predict`
(BaseStableDiffusionXLSetup.py:179-358), `BaseFluxSetup.predict` (BaseFluxSetup.py:193-375) and their
`calculate_loss` are executed from baseline/_ref (byte-for-byte copies, tools/make_baseline_ref.py) by
baseline/ref_callers.py:

* CPU (not gpu): the unmodified callers on the reference's own mixins reproduce, bit for bit, the restated callers that
  baseline/ref_arm.py and the harness use everywhere else — so every parity test written against the restatement is a
  test against the real caller;
* GPU: the SAME caller source with the seam modules redirected to onetrainer_b200 — the
  reference's `predict()` / `calculate_loss()` calling this repo's `_create_noise`, `_get_timestep_discrete`,
  `_add_noise_discrete`, `_diffusion_losses` / `_flow_matching_losses` with the reference's own TrainConfig and model
  objects, the network hooked by this repo's LoRAModuleWrapper — against the reference arm fed the same draws: loss
  within 1e-3 relative, prediction and LoRA gradients to bf16 accuracy.
"""
import math

import pytest
import torch


def _have_callers():
    try:
        from baseline import ref_callers
        return ref_callers.available()
    except Exception:  # noqa: BLE001
        return False


needs_ref = pytest.mark.skipif(not _have_callers(), reason="baseline/_ref (with the family callers) did not travel")


def _network(kind, device, dtype):
    from onetrainer_b200.harness import standin
    torch.manual_seed(1234)
    with torch.device(device):
        net = {"sdxl": standin.sdxl_unet, "sd15": standin.sd15_unet, "tiny": standin.tiny_unet, "tiny_sd": standin.tiny_unet,
               "flux": standin.flux_transformer, "tiny_flux": standin.tiny_flux}[kind]()
    return net.to(dtype=dtype).requires_grad_(False).train()


def _scheduler(kind, device, prediction_type="epsilon"):
    from onetrainer_b200.harness import standin, trainer
    if "flux" in kind:
        return standin.StandInFlowScheduler(1000, device)
    return standin.StandInScheduler(trainer.standin_betas().to(device), prediction_type)


def _batch(kind, B, device, latent_hw=None, masked=False):
    from baseline import ref_callers
    from onetrainer_b200.harness import trainer
    base_kind = {"tiny_sd": "tiny"}.get(kind, kind)
    b = trainer.synthetic_batch(base_kind, B, device, latent_hw=latent_hw, masked=masked)
    if kind == "tiny":                                   # the tiny stand-in driven through the SDXL caller
        px = float(b["latent_image"].shape[-1] * 8)
        b["text_encoder_2_pooled_state"] = torch.zeros(B, 8)
        b["original_resolution"] = [torch.full((B,), px), torch.full((B,), px)]
        b["crop_offset"] = [torch.zeros(B), torch.zeros(B)]
        b["crop_resolution"] = [torch.full((B,), px), torch.full((B,), px)]
    b = trainer.batch_to_device(b, device, non_blocking=False)
    return ref_callers.caller_batch("sd15" if kind == "tiny_sd" else kind, b)


def _setup_for(ns, kind, device):
    cls = ns.FluxSetup if "flux" in kind else ns.SDXLSetup if kind in ("sdxl", "tiny") else ns.SDSetup
    return cls(torch.device(device), torch.device("cpu"), False)


# ------------------------------------------------------------------------------------------------ CPU
@needs_ref
@pytest.mark.parametrize("kind,ptype", [("tiny", "epsilon"), ("tiny", "v_prediction"), ("tiny_sd", "epsilon"), ("tiny_flux", "flow")])
def test_restated_callers_equal_the_unmodified_callers_on_cpu(kind, ptype):
    """fp32 on the CPU (BASELINE.json configs[0] flavour): the real predict()/calculate_loss() and the restatement used by
    ref_arm / the harness give the SAME bits (same torch generator seed, same reference mixins underneath)."""
    from baseline import ref_arm, ref_callers
    ns = ref_callers.load(dropin=False)
    R = ns.R
    assert all(m.startswith("modules.") for m in ns.mixin_owner.values()), ns.mixin_owner
    flux = "flux" in kind
    rkind = "tiny" if kind == "tiny_sd" else kind
    cfg = ref_callers.caller_config(ns, "sd15" if kind == "tiny_sd" else kind, train_device="cpu", batch_size=2, lora_rank=8, lora_alpha=1.0,
                                   loss_weight_fn=R.LossWeight.CONSTANT if flux else R.LossWeight.MIN_SNR_GAMMA,
                                   timestep_distribution=R.TimestepDistribution.LOGIT_NORMAL if flux else R.TimestepDistribution.UNIFORM)
    # the restated arm (ref_arm): builds the stand-in, hooks the reference wrapper
    rmodel, rsetup = ref_arm.build_reference_model(rkind, cfg, "cpu", train_dtype=torch.float32, prediction_type=ptype, lora_up_std=0.02)
    rmodel.sdxl = kind == "tiny"
    if kind == "tiny":
        rmodel.vae_scaling_factor = 0.13025                # the SDXL VAE's, as make_model sets it for this caller
    batch = _batch(kind, 2, "cpu", latent_hw=16)
    data_r = rsetup.predict(rmodel, batch, cfg, rmodel.train_progress)
    loss_r = rsetup.calculate_loss(rmodel, batch, data_r, cfg)
    # the unmodified caller on the same hooked network
    model = ref_callers.make_model(ns, "sd15" if kind == "tiny_sd" else kind, rmodel.unet, rmodel.noise_scheduler, torch.float32, "cpu",
                                   lora_wrapper=rmodel.unet_lora)
    setup = _setup_for(ns, "sd15" if kind == "tiny_sd" else kind, "cpu")
    data = setup.predict(model, batch, cfg, model.train_progress)
    loss = setup.calculate_loss(model, batch, data, cfg)
    assert torch.equal(data["timestep"], data_r["timestep"])
    assert torch.equal(data["target"], data_r["target"])
    assert torch.equal(data["predicted"], data_r["predicted"])
    assert torch.equal(loss, loss_r)
    if not flux:
        assert data["prediction_type"] == data_r["prediction_type"] == ptype


@needs_ref
def test_dropin_arm_executes_the_same_caller_files_with_only_the_seams_redirected():
    """Host logic of the drop-in arm (no GPU): the three family-caller source files are executed a second time from the very
    same paths, the four seam mixins in their MRO come from onetrainer_b200, everything else in the MRO (BaseModelSetup, the
    debug / embedding mixins) is still the reference's, and the reference arm itself stays untouched by the redirection."""
    from baseline import ref_callers
    ref = ref_callers.load(dropin=False)
    own = ref_callers.load(dropin=True)
    assert own.caller_files == ref.caller_files and all("baseline/_ref/modules/modelSetup/Base" in f for f in own.caller_files.values())
    seams = {n.rsplit(".", 1)[-1] for n in ref_callers.SEAM_MODULES}
    for ns_ in (own.SDXLSetup, own.SDSetup, own.FluxSetup):
        owners = {c.__name__: c.__module__ for c in ns_.__mro__ if c.__name__.startswith("ModelSetup")}
        assert {k for k, m in owners.items() if m.startswith("onetrainer_b200.")} == set(owners) & seams, owners
        assert any(c.__module__ == "modules.modelSetup.BaseModelSetup" for c in ns_.__mro__)
    for ns_ in (ref.SDXLSetup, ref.SDSetup, ref.FluxSetup):
        assert all(c.__module__.startswith("modules.") for c in ns_.__mro__ if c.__name__.startswith("ModelSetup"))
    # the predict() / calculate_loss() the two arms run are the same functions' source
    import inspect
    for a, b in ((own.SDXLSetup, ref.SDXLSetup), (own.SDSetup, ref.SDSetup), (own.FluxSetup, ref.FluxSetup)):
        assert inspect.getsource(a.predict) == inspect.getsource(b.predict)
        assert inspect.getsource(a.calculate_loss) == inspect.getsource(b.calculate_loss)
    import sys
    assert sys.modules["modules.modelSetup.mixin.ModelSetupNoiseMixin"].__file__.endswith("baseline/_ref/modules/modelSetup/mixin/ModelSetupNoiseMixin.py")


# ------------------------------------------------------------------------------------------------ GPU
def _run_dropin_vs_reference(dev, kind, B, rank, cfg_kw, ptype="epsilon", latent_hw=None, masked=False):
    from baseline import ref_arm, ref_callers
    from onetrainer_b200.modules.module.LoRAModule import LoRAModuleWrapper
    ref = ref_callers.load(dropin=False)
    own = ref_callers.load(dropin=True)
    R = ref.R
    seams = {n.rsplit(".", 1)[-1] for n in ref_callers.SEAM_MODULES}
    assert all(m.startswith("onetrainer_b200.") for k, m in own.mixin_owner.items() if k in seams), own.mixin_owner
    assert all(m.startswith("modules.") for k, m in ref.mixin_owner.items()), ref.mixin_owner
    assert own.caller_files == ref.caller_files          # the very same source files
    flux = "flux" in kind
    ckind = "sd15" if kind == "tiny_sd" else kind
    cfg = ref_callers.caller_config(ref, ckind, train_device=str(dev), batch_size=B, lora_rank=rank, lora_alpha=1.0, **{k: v(R) if callable(v) else v for k, v in cfg_kw.items()})
    net = _network(kind, dev, torch.bfloat16)
    sched = _scheduler(kind, dev, ptype)
    batch = _batch(kind, B, dev, latent_hw=latent_hw, masked=masked)
    filt = ["attn", "ff.net"] if flux else ["attentions"]

    # ---- drop-in arm: this repo's wrapper + mixins under the reference's unmodified caller, config and model classes
    wrap = LoRAModuleWrapper(net, "lora_transformer" if flux else "lora_unet", cfg, filt)
    wrap.to(dtype=torch.float32)
    g = torch.Generator(device=dev).manual_seed(77)
    for m in wrap.lora_modules.values():
        m.lora_up.weight.data.normal_(0.0, 0.02, generator=g)
    wrap.hook_to_module()
    wrap.requires_grad_(True)
    model = ref_callers.make_model(own, ckind, net, sched, torch.bfloat16, dev, lora_wrapper=wrap)
    setup = _setup_for(own, ckind, dev)
    draws = {}
    noise_fn, ts_fn = setup._create_noise, setup._get_timestep_discrete
    setup._create_noise = lambda *a, **k: draws.setdefault("noise", noise_fn(*a, **k))
    setup._get_timestep_discrete = lambda *a, **k: draws.setdefault("timestep", ts_fn(*a, **k))
    data = setup.predict(model, batch, cfg, model.train_progress)
    loss = setup.calculate_loss(model, batch, data, cfg)
    loss.backward()
    torch.cuda.synchronize()
    mine = {k: (m.lora_down.weight.grad.clone(), m.lora_up.weight.grad.clone()) for k, m in wrap.lora_modules.items()}
    pred_mine, loss_mine = data["predicted"].detach().float().clone(), loss.item()
    target_mine = data["target"].detach().clone()
    wrap.remove_hook_from_module()

    # ---- reference arm: everything unmodified, fed the same draws
    rwrap = R.LoRAModuleWrapper(net, "lora_transformer" if flux else "lora_unet", cfg, filt)
    rwrap.to(dtype=torch.float32)
    with torch.no_grad():
        for k, m in wrap.lora_modules.items():
            rwrap.lora_modules[k].lora_down.weight.copy_(m.lora_down.weight)
            rwrap.lora_modules[k].lora_up.weight.copy_(m.lora_up.weight)
    rwrap.hook_to_module()
    rwrap.requires_grad_(True)
    rmodel = ref_callers.make_model(ref, ckind, net, sched, torch.bfloat16, dev, lora_wrapper=rwrap)
    rsetup = _setup_for(ref, ckind, dev)
    rsetup._create_noise = lambda src, *a, **k: draws["noise"].to(src.dtype)
    rsetup._get_timestep_discrete = lambda *a, **k: draws["timestep"]
    rdata = rsetup.predict(rmodel, batch, cfg, rmodel.train_progress)
    rloss = rsetup.calculate_loss(rmodel, batch, rdata, cfg)
    rloss.backward()
    torch.cuda.synchronize()
    assert torch.equal(rdata["timestep"], data["timestep"])
    assert torch.equal(rdata["target"], target_mine), "target (epsilon / velocity / flow) must be bit-identical given the same draws"
    if not flux:
        assert data["prediction_type"] == rdata["prediction_type"] == ptype
    pr = rdata["predicted"].detach().float()
    rel = (pred_mine - pr).norm() / pr.norm()
    assert rel < 2e-2, f"prediction differs by {rel:.3e} of its norm"
    assert loss_mine == pytest.approx(rloss.item(), rel=1e-3), (loss_mine, rloss.item())
    cosines = []
    for k, m in rwrap.lora_modules.items():
        for a, b in ((mine[k][0], m.lora_down.weight.grad), (mine[k][1], m.lora_up.weight.grad)):
            a, b = a.double().flatten(), b.double().flatten()
            cosines.append(float(a @ b) / max(math.sqrt(float(a @ a) * float(b @ b)), 1e-300))
    cosines.sort()
    assert cosines[len(cosines) // 2] >= 0.995 and cosines[0] >= 0.9, (cosines[len(cosines) // 2], cosines[0])
    rwrap.remove_hook_from_module()
    return loss_mine, rloss.item()


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("kind,ptype", [("tiny", "epsilon"), ("tiny", "v_prediction"), ("tiny_sd", "epsilon"), ("tiny_flux", "flow")])
def test_unmodified_callers_on_the_dropin_modules_tiny(dev, kind, ptype):
    flux = "flux" in kind
    kw = {"loss_weight_fn": (lambda R: R.LossWeight.CONSTANT) if flux else (lambda R: R.LossWeight.MIN_SNR_GAMMA),
          "timestep_distribution": (lambda R: R.TimestepDistribution.LOGIT_NORMAL) if flux else (lambda R: R.TimestepDistribution.UNIFORM)}
    _run_dropin_vs_reference(dev, kind, 2, 8, kw, ptype=ptype, latent_hw=16)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C2_sd15_b8_masked", "C3_sdxl_b4_minsnr", "C5_flux_b1_flow"])
def test_unmodified_callers_on_the_dropin_modules_at_config_size(dev, name):
    """BASELINE.json configs 2, 3 and 5 through the reference's unmodified predict()/calculate_loss()."""
    torch.cuda.empty_cache()
    if name.startswith("C2"):
        _run_dropin_vs_reference(dev, "sd15", 8, 16, {"masked_training": True, "normalize_masked_area_loss": True, "unmasked_weight": 0.1}, masked=True)
    elif name.startswith("C3"):
        _run_dropin_vs_reference(dev, "sdxl", 4, 32, {"loss_weight_fn": lambda R: R.LossWeight.MIN_SNR_GAMMA})
    else:
        _run_dropin_vs_reference(dev, "flux", 1, 16, {"timestep_distribution": lambda R: R.TimestepDistribution.LOGIT_NORMAL})
    torch.cuda.empty_cache()
