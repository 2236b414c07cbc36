"""pytest configuration: import paths, the ``gpu`` marker, shared session fixtures."""

import json
import os
import sys
import warnings

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(REPO, "qwen-megakernel-tts_b200")
for p in (REPO, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")
warnings.filterwarnings("ignore", category=UserWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: long-running")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def bf16_from_bits(bits: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(bits).astype(np.int16)).view(torch.bfloat16)


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN, "meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz"))
            for name in ("talker_config1", "talker_mixed", "cp_config2")}


@pytest.fixture(scope="session")
def cpu_weights(golden_meta):
    """The seeded synthetic checkpoint (CPU); must regenerate bit-identically on every host."""
    from qwen_megakernel.synthetic import synthetic_tts_weights, weights_fingerprint
    torch.set_num_threads(os.cpu_count() or 1)
    w = synthetic_tts_weights(seed=golden_meta["seed_weights"])
    assert weights_fingerprint(w) == golden_meta["weights_fingerprint"], \
        "synthetic weights differ from the ones the golden fixtures were generated with"
    return w


@pytest.fixture(scope="session")
def gpu_weights(cpu_weights):
    from qwen_megakernel.synthetic import weights_to
    return weights_to(cpu_weights, "cuda")
