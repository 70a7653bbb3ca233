import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def hostmodel():
    sys.path.insert(0, str(ROOT / "tests"))
    from hostmodel import model
    return model


@pytest.fixture(scope="session")
def codec():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from datacompressionfloat_b200 import Codec
    c = Codec(0)
    yield c
    c.close()


def synth_words(kind: str, n_data_words: int, seed: int = 0) -> np.ndarray:
    """256 MRC header words + n_data_words float32 words of distribution `kind` (small, fast)."""
    from datacompressionfloat_b200 import synth
    rng = np.random.default_rng(seed + 17)
    if kind == "G":
        d = rng.standard_normal(n_data_words, dtype=np.float32)
    elif kind == "P":
        d = rng.poisson(2.0, n_data_words).astype(np.float32)
    elif kind == "S":
        x = np.linspace(0, 40 * np.pi, n_data_words, dtype=np.float32)
        d = (np.sin(x) * np.cos(x / 7) + rng.normal(0, 0.25, n_data_words)).astype(np.float32)
    elif kind == "Z":
        d = np.zeros(n_data_words, dtype=np.float32)
    elif kind == "R":  # random bits: every plane incompressible
        d = rng.integers(0, 2**32, n_data_words, dtype=np.uint64).astype(np.uint32).view(np.float32)
    else:
        raise ValueError(kind)
    hdr = synth.mrc_header(max(n_data_words, 1), 1, 1)
    hdr[10:200] = rng.integers(0, 2**32, 190, dtype=np.uint64).astype(np.uint32)  # header words must survive unmasked
    return np.concatenate([hdr, d.view(np.uint32)])
