import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # parity tests inspect the cross-modal block's token-level output, which the fused kernel keeps on-chip in production
    os.environ.setdefault("ROBOVLN_KEEP_TOKENS", "1")
    # the torch references of the GPU tests must be true fp32 (no TF32 convs / matmuls)
    import torch

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
