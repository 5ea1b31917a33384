import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "detecting-convolutional-codes-via-markovian-statistics_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

CODES = {
    "c75": dict(k=1, n=2, m=2, gen=[[[1, 1, 1]], [[1, 0, 1]]]),
    "c65": dict(k=1, n=2, m=2, gen=[[[1, 1, 0]], [[1, 0, 1]]]),
    "m3a": dict(k=1, n=2, m=3, gen=[[[1, 1, 1, 1]], [[1, 0, 1, 1]]]),
    "m3b": dict(k=1, n=2, m=3, gen=[[[1, 0, 1, 1]], [[1, 1, 1, 1]]]),
    "r13": dict(k=1, n=3, m=2, gen=[[[1, 1, 1]], [[1, 0, 1]], [[1, 1, 0]]]),
    "m1": dict(k=1, n=2, m=1, gen=[[[1, 1]], [[1, 0]]]),
    "m4a": dict(k=1, n=2, m=4, gen=[[[1, 1, 0, 0, 1]], [[1, 1, 0, 1, 1]]]),     # (31,33): S = 25 751
    "m4b": dict(k=1, n=2, m=4, gen=[[[1, 1, 0, 1, 1]], [[1, 1, 0, 0, 1]]]),     # outputs swapped
    "m4c": dict(k=1, n=2, m=4, gen=[[[1, 0, 0, 1, 1]], [[1, 1, 1, 0, 1]]]),     # (23,35): S = 150 743
    "m5": dict(k=1, n=2, m=5, gen=[[[1, 0, 1, 0, 1, 1]], [[1, 1, 1, 1, 0, 1]]]),            # (53,75)
    "m6": dict(k=1, n=2, m=6, gen=[[[1, 0, 1, 1, 0, 1, 1]], [[1, 1, 1, 1, 0, 0, 1]]]),      # (133,171)
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_unavailable_reason():
    """None when libmvd.so loads and mvd_create succeeds on device 0, else why not."""
    import ctypes as C
    try:
        from mvd import _capi
        lib = _capi.load()
    except Exception as exc:                       # missing / stale library
        return f"libmvd.so unusable: {exc}"
    ctx = C.c_void_p()
    rc = lib.mvd_create(C.byref(ctx), 0)
    if rc != 0:
        return "no CUDA device: " + (lib.mvd_last_error(None) or b"").decode(errors="replace")
    lib.mvd_destroy(ctx)
    return None


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU-only box: gpu-marked tests are skipped (with the reason), not failed, so host-logic
    regressions are not buried.  When the marker expression asks for gpu tests explicitly (-m gpu) nothing is
    skipped: on the GPU box a missing device or library must fail loudly."""
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for it in gpu_items:
        it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    out = {}
    for name in ("code_kats", "sim_kats", "experiments", "m4_kats", "alpha_kats", "parity_kats"):
        with open(os.path.join(GOLDEN, name + ".json")) as f:
            out[name] = json.load(f)
    return out


@pytest.fixture(scope="session")
def codes_spec():
    return CODES
