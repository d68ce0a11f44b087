"""Ship the UNMODIFIED reference files of the hot path to the GPU box: /root/reference -> baseline/_ref/.

`/root/reference` exists only in the build container; `baseline/_ref/` is git-ignored (reference sources never enter
this repository's history) but NOT gpurun-ignored, so it travels with the snapshot like the built `.so` files.  What is
copied is exactly the set of reference files Python loads when the hot-path modules of SURVEY.md §8(a) are imported
(the import closure, found by importing them here with the two shims of SURVEY.md F7) — byte for byte, same relative
paths.  `baseline/ref_arm.py` (tracked, ours) then drives them: LoRAModuleWrapper, the noise / add-noise / loss mixins,
patch_adamw (+ stochastic rounding), EMAModuleWrapper, TrainConfig.

    python tools/make_baseline_ref.py            # idempotent; __graft_entry__.build() calls it when /root/reference exists
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUT = REPO / "baseline" / "_ref"

# the entry modules of the path; everything they import from the reference comes along
ENTRY = [
    "modules.module.LoRAModule", "modules.module.EMAModule",
    "modules.modelSetup.mixin.ModelSetupNoiseMixin", "modules.modelSetup.mixin.ModelSetupDiffusionMixin",
    "modules.modelSetup.mixin.ModelSetupFlowMatchingMixin", "modules.modelSetup.mixin.ModelSetupDiffusionLossMixin",
    "modules.util.loss.masked_loss", "modules.util.DiffusionScheduleCoefficients", "modules.util.bf16_stochastic_rounding",
    "modules.util.optimizer.adamw_extensions", "modules.util.optimizer.adam_extensions",
    "modules.util.optimizer.adafactor_extensions", "modules.util.optimizer.CAME",
    "modules.util.NamedParameterGroup", "modules.util.config.TrainConfig",
    "modules.util.convert.rescale_noise_scheduler_to_zero_terminal_snr",
]
# the family callers of SURVEY.md §8 row A16 (predict / calculate_loss of the three Base*Setup classes), run unmodified by
# baseline/ref_callers.py with a placeholder `diffusers` package
CALLERS = ["modules.modelSetup.BaseStableDiffusionSetup", "modules.modelSetup.BaseStableDiffusionXLSetup", "modules.modelSetup.BaseFluxSetup"]


def closure() -> list[Path]:
    sys.path.insert(0, str(REPO))
    from baseline import ref_arm, ref_callers
    ref_arm.install_shims()
    sys.path.insert(0, str(REF))
    import importlib
    for name in ENTRY:
        importlib.import_module(name)
    ref_callers.install_caller_shims()
    for name in CALLERS:
        importlib.import_module(name)
    files = set()
    for mod in list(sys.modules.values()):
        f = getattr(mod, "__file__", None)
        if f and Path(f).resolve().is_relative_to(REF):
            files.add(Path(f).resolve())
    return sorted(files)


def make(verbose: bool = True) -> Path | None:
    if not REF.exists():
        if verbose:
            print(f"{REF} is not present (GPU box): keeping the prebuilt {OUT}")
        return OUT if OUT.exists() else None
    import subprocess
    # the closure is computed in a fresh interpreter so that importing the reference never pollutes the caller
    res = subprocess.run([sys.executable, __file__, "--list"], check=True, capture_output=True, text=True)
    files = [Path(line) for line in res.stdout.splitlines() if line.startswith(str(REF))]
    if OUT.exists():
        shutil.rmtree(OUT)
    manifest = {}
    for f in files:
        rel = f.relative_to(REF)
        dst = OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(f, dst)
        manifest[str(rel)] = hashlib.sha256(f.read_bytes()).hexdigest()
    (OUT / "MANIFEST.json").write_text(json.dumps({"source": str(REF), "files": manifest}, indent=1, sort_keys=True))
    if verbose:
        print(f"copied {len(files)} reference files ({sum(f.stat().st_size for f in files) / 1e3:.0f} kB) to {OUT}")
    return OUT


if __name__ == "__main__":
    if "--list" in sys.argv:
        for f in closure():
            print(f)
    else:
        make()
