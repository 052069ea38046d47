import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests" / "golden" / "objectives.json").read_text())


@pytest.fixture(scope="session")
def built():
    """Build (if stale) and return the native artefacts; tests never fall back to anything else."""
    from miplib_b200 import build
    build.build_all()
    return build
