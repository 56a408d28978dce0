"""The product library loads on a CPU-only box and exports every symbol include/oduck.h declares (no CUDA call is made)."""
import ctypes as C
import os
import re

import pytest

from open_duck_playground_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "oduck.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(oduck_[a-z_]+)\s*\(", text)))


def test_header_declares_the_surface():
    names = _declared()
    for n in ("oduck_create", "oduck_reset", "oduck_step", "oduck_physics_substeps", "oduck_randomize", "oduck_policy_forward",
              "oduck_get_buffer", "oduck_set_state", "oduck_last_error"):
        assert n in names


def test_cuda_library_exports_every_declared_symbol():
    path = capi.cuda_library_path()
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = capi.Library(path, is_device=True)          # dlopen + ABI version + sizeof(struct) checks
    for n in _declared():
        assert hasattr(lib.lib, n), n


def test_oracle_exports_the_same_surface(oracle):
    for n in _declared():
        assert hasattr(oracle.lib, n), n


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(capi.OduckError):
        capi.Library(str(tmp_path / "liboduck_cuda.so"), is_device=True)


def test_struct_sizes_match(oracle):
    assert oracle.lib.oduck_sizeof_model() == C.sizeof(capi.OduckModel)
    assert oracle.lib.oduck_sizeof_env_config() == C.sizeof(capi.OduckEnvConfig)
