"""A fixed slice of tools/fuzz_parity.py in the GPU suite: random batches (scoring scheme, gap
penalty, mode, lengths on the kernels' seams, route, bands) through the C ABI against the oracle,
bit for bit.  The long runs are made with the tool itself (profiles/r01_fuzz_parity.json)."""
import importlib.util
import os
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

_TOOL = Path(__file__).resolve().parent.parent / "tools" / "fuzz_parity.py"


def _load_tool():
    spec = importlib.util.spec_from_file_location("fuzz_parity", _TOOL)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


@pytest.mark.parametrize("first_seed", [9000, 9040, 9080])
def test_fuzz_slice(first_seed):
    fuzz = _load_tool()
    saved = os.environ.get("B200ALIGN_ROUTE")
    try:
        for seed in range(first_seed, first_seed + 40):
            info, err = fuzz.iteration(seed)
            assert err is None, (info, err)
    finally:
        if saved is None:
            os.environ.pop("B200ALIGN_ROUTE", None)
        else:
            os.environ["B200ALIGN_ROUTE"] = saved
