"""The stand-alone gradient-check drivers (``tools/gradient_check.py``; counterparts of the reference's
``tests/gradient_test_{torch,distdl,distdl_bcast,dfno}.py``) under pytest: gloo, 2 ranks."""
import importlib.util
import os

import pytest

from dfno_b200.utils.testing import run_distributed

_TOOL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "gradient_check.py")


def _load():
    spec = importlib.util.spec_from_file_location("gradient_check", _TOOL)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _case(rank, ws, case):
    return _load().run_case(case, verbose=False)


def test_pure_torch_control():
    assert _load().run_case("torch", verbose=False) == []


@pytest.mark.parametrize("case", ["transpose", "transpose-linear", "bcast", "dfno"])
def test_distributed_driver(case):
    """``transpose-linear`` is the composition the reference reports as failing its own check."""
    res = run_distributed(_case, 2, case, timeout=600)
    assert all(not bad for bad in res), "\n".join(sum(res, []))
