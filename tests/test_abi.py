"""The C-ABI library loads without a GPU and exports exactly what include/otb200.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (REPO / "include" / "otb200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(otb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for required in ("otb_lora_linear", "otb_lora_grad_ab", "otb_noise_addnoise", "otb_loss_fwd_bwd",
                     "otb_adamw_sr_ema_multi", "otb_ema_multi", "otb_grad_norm_multi", "otb_last_error"):
        assert required in syms


def test_library_exports_every_declared_symbol():
    from onetrainer_b200 import _native as N
    lib = N.lib()
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/otb200.h but not exported: {missing}"
    assert set(N.SYMBOLS) == set(_declared_symbols()), "ctypes binding and header disagree on the symbol list"


def test_struct_layouts_match_the_header():
    from onetrainer_b200 import _native as N
    sizes = (C.c_int32 * len(N._STRUCTS))()
    assert N.lib().otb_struct_sizes(sizes) == 0
    assert [int(s) for s in sizes] == [C.sizeof(s) for s in N._STRUCTS]
    assert N.lib().otb_abi_version() == N.ABI_VERSION


def test_flavor_constants_agree_across_header_binding_and_oracle():
    from onetrainer_b200 import _native as N
    from oracle import otb_oracle as O
    text = (REPO / "include" / "otb200.h").read_text()
    for name in ("DIV_RECIPROCAL", "ADDCDIV_MUL_LAST", "FMA_LERP", "FMA_ADDCMUL", "FMA_ADDCDIV", "FMA_ADD_ALPHA",
                 "ADDCMUL_SQUARE_FIRST", "ALPHA_IN_DTYPE"):
        v = int(re.search(rf"#define OTB_FLAVOR_{name}\s+(\d+)u", text).group(1))
        assert getattr(N, f"FLAVOR_{name}") == v == getattr(O, f"FL_{name}")
    assert N.FLAVOR_TORCH_CPU == O.FLAVOR_TORCH_CPU and N.FLAVOR_TORCH_CUDA == O.FLAVOR_TORCH_CUDA


def test_argument_errors_surface_as_python_exceptions():
    """No compute without a GPU — but argument validation happens on the host before any launch."""
    from onetrainer_b200 import _native as N
    a = N.LoraLinearArgs(M=0, N=8, K=8)
    rc = N.lib().otb_lora_linear(C.byref(a), None)
    assert rc == N.OTB_EINVAL
    with pytest.raises(ValueError):
        N.check(rc, "otb_lora_linear")
    a = N.LoraLinearArgs(x=16, w=16, y=16, down=16, up=16, M=128, N=64, K=64, R=12)   # rank not a multiple of 8
    rc = N.lib().otb_lora_linear(C.byref(a), None)
    assert rc == N.OTB_EUNSUPPORTED
    with pytest.raises(NotImplementedError):
        N.check(rc, "otb_lora_linear")


def test_stochastic_round_argument_validation():
    """otb_stochastic_round (bf16_stochastic_rounding.py as a launch): an empty tensor is a no-op, bad arguments are errors —
    checked on the host before anything is launched."""
    from onetrainer_b200 import _native as N
    lib = N.lib()
    assert lib.otb_stochastic_round(C.byref(N.SrArgs(n=0, mode=N.SR_COPY)), None) == N.OTB_OK
    ok = dict(target=16, src=32, n=64, mode=N.SR_ADD, src_dtype=N.F32)
    for bad in (dict(n=-1), dict(target=None), dict(src=None), dict(mode=7), dict(mode=N.SR_ADDCDIV),      # addcdiv without tensor2
                dict(mode=N.SR_COPY, src_dtype=N.BF16), dict(rng_base=4), dict(rng_base=-8), dict(src=34)):
        assert lib.otb_stochastic_round(C.byref(N.SrArgs(**{**ok, **bad})), None) == N.OTB_EINVAL, bad
    assert lib.otb_stochastic_round(C.byref(N.SrArgs(**{**ok, "src_dtype": N.F16})), None) == N.OTB_EUNSUPPORTED


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from onetrainer_b200 import _native as N
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", tmp_path / "libotb200.so")
    with pytest.raises(RuntimeError, match="no non-CUDA fallback"):
        N.lib()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under onetrainer_b200/ may reference it."""
    for f in (REPO / "onetrainer_b200").rglob("*.py"):
        src = f.read_text()
        assert "otb_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f
    for f in (REPO / "onetrainer_b200" / "csrc").glob("*"):
        src = f.read_text()
        assert "otb_o_" not in src, f"{f} calls an oracle symbol"
        assert not [ln for ln in src.splitlines() if ln.strip().startswith("#include") and "oracle" in ln], f
