"""The C-ABI library loads without a GPU, exports every symbol include/pvae_sm100.h declares, and fails loudly (no CPU
fallback) when asked to compute without a device.  CPU only: no compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

from physicsvae_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pvae_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pvae_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_abi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    for name in _declared():
        assert getattr(lib, name) is not None
    assert lib.pvae_abi_version() == _abi.PVAE_ABI_VERSION
    assert isinstance(_abi.launch_count(), int)


def test_struct_layout_matches_header():
    # pvae_net_desc: int32 + 8 int32 + 8 int32 + 2 int32 (input override); pvae_model_desc: 6 int32 + 4 nets
    assert C.sizeof(_abi.NetDesc) == 4 * 19
    assert C.sizeof(_abi.ModelDesc) == 4 * 6 + 4 * C.sizeof(_abi.NetDesc)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _abi.load()
    h = C.c_void_p(0)
    d = _abi.ModelDesc()
    d.dim_state_body, d.dim_action, d.latent_dim, d.max_batch, d.precision = 3, 2, 2, 4, 1
    rc = lib.pvae_create(C.byref(h), C.byref(d), 0)
    assert rc == -2 and b"no CPU fallback" in lib.pvae_last_error()
    with pytest.raises(_abi.PvaeError):
        _abi.check(rc)
    from physicsvae_b200.engine import Engine
    with pytest.raises(_abi.PvaeError):
        Engine(3, 2, 2, {})
