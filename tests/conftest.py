import json, os, sys, pathlib
import pytest

# Several ranks of a commit group share ONE device in tests/test_gpu_shard_group.py: their streams must not be multiplexed onto the same
# hardware queue (a rank spinning in a flag barrier would hold back the peer it waits for).  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests" / "golden" / "sm_all_proof.json").read_text())
