import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # tests fill every model with synthetic weights; the loud "text tower is randomly initialised" warning
    # (fiber_b200/modules/roberta.py: from_pretrained without local weights) is exercised by its own test
    config.addinivalue_line("filterwarnings", "ignore:fiber_b200. RobertaModel.from_pretrained:RuntimeWarning")


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fiber_b200 import lib
    l = lib.load()
    lib.check(l.fiber_init(), "init")
    return torch.device("cuda:0")


@pytest.fixture
def unfused_mlm_ce():
    """The reference's logits -> F.cross_entropy form of the MLM loss (tests that compare mlm_logits element-wise); the
    default on the GPU is the fused decoder + cross-entropy, which never materialises them."""
    from fiber_b200.modules import objectives as OBJ
    old = OBJ.FUSED_MLM_CE
    OBJ.set_fused_mlm_ce(False)
    yield
    OBJ.set_fused_mlm_ce(old)
