import glob
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden_cases():
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.json"))):
        if os.path.basename(p).startswith(("counts_", "path_", "shapes_")):  # counting stage / PathAligner fixtures (own tests)
            continue
        with open(p) as f:
            out.append(json.load(f))
    return out


def strip_status(results):
    """capi/emulator dicts carry a status field the reference has no counterpart for: it must be 0."""
    out = []
    for r in results:
        r = dict(r)
        assert r.pop("status", 0) == 0, r
        r.pop("clipped", None)  # checked separately against the reference's read filters (test_read_filters.py)
        out.append(r)
    return out


@pytest.fixture(scope="session")
def built():
    """Build the native pieces once per session (oracle always; CUDA library if nvcc is around)."""
    import __graft_entry__ as g
    g.build()
    return True


def ref_graph_shapes():
    """The reference's own graph JSONs as committed shapes (tools/make_golden_graphs.py)."""
    with open(os.path.join(GOLDEN_DIR, "shapes_ref_graphs.json")) as f:
        return json.load(f)["graphs"]
