import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_usable() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Plain ``pytest`` on a box without a usable CUDA device skips the ``gpu`` cases instead of failing them one by one
    (the product has no CPU fallback: the library raises).  On a GPU box nothing is skipped."""
    if _cuda_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (the gpu-marked tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The fp64 CPU oracle library (built on demand)."""
    from oracle import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def model_backlash():
    from open_duck_playground_b200 import constants
    from open_duck_playground_b200.mjcf import CompiledModel
    return CompiledModel.load(constants.task_to_blob("flat_terrain_backlash"))


@pytest.fixture(scope="session")
def poly_table():
    from open_duck_playground_b200 import constants
    from open_duck_playground_b200.poly_reference_motion import PolyTable
    return PolyTable.load(constants.POLY_BLOB)


def make_handle(lib, model, poly, n, **cfg_kw):
    from open_duck_playground_b200 import capi, config
    cfg = config.default_config()
    overrides = cfg_kw.pop("overrides", None)
    if overrides:
        cfg.update_from_flattened_dict(overrides)
    ms = capi.model_to_struct(model)
    cs, keep = config.build_env_config(model, cfg, poly, **cfg_kw)
    h = lib.create(ms, cs, n)
    h._keep = keep
    return h
