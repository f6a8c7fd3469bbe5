import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_meta():
    return json.loads((GOLDEN / "meta.json").read_text())


def load_pcm_case(path):
    """golden npz -> (x float32 exactly as the generating script built it, npz dict)."""
    d = np.load(path)
    x = (d["pcm"].astype(np.float32) * np.float32(1.0 / 32768.0)) * np.float32(d["gain"])
    return x, d


@pytest.fixture(scope="session")
def standin_encoder():
    from amphibian_vae_latent_detector_b200.encoder import build_standin_encoder
    return build_standin_encoder(seed=123)


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine3s(standin_encoder):
    """Engine for 3 s chunks with the stand-in encoder loaded (GPU tests only)."""
    if not _cuda_ok():
        pytest.skip("no CUDA device")
    from amphibian_vae_latent_detector_b200.engine import Engine
    eng = Engine(0, chunk_len=144000, max_batch=64)
    eng.load_encoder(standin_encoder)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def engine5s():
    if not _cuda_ok():
        pytest.skip("no CUDA device")
    from amphibian_vae_latent_detector_b200.engine import Engine
    eng = Engine(0, chunk_len=240000, max_batch=16)
    yield eng
    eng.close()
