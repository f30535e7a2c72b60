import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Reference outputs frozen by oracle/make_golden.py: (manifest dict, npz arrays)."""
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as f:
        manifest = json.load(f)
    arrays = np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz"))
    return manifest, arrays


def case_table():
    from oracle import golden_cases
    return ({c[0]: c for c in golden_cases.CASES},
            {c[0]: c for c in golden_cases.ERROR_CASES},
            {c[0]: c for c in golden_cases.SPM_CASES})
