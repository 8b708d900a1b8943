import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def cuda_device():
    import torch
    if os.environ.get('SFB_TEST_DRY_RUN') == '1' and not torch.cuda.is_available():
        # developer aid, never set by the driver: run the *test logic* of tests/test_train_gpu.py on the CPU stand-ins of
        # tests/fake_ops.py (shapes, tolerances, oracle plumbing), so that on the GPU box only the kernels themselves can fail
        import fake_ops
        mp = pytest.MonkeyPatch()
        fake_ops.install(mp, round_bf16=True, names=fake_ops.ALL + fake_ops.ENCODER_FWD + fake_ops.N1_BWD)
        return torch.device('cpu')
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from synchformer_b200 import ops
    ops.device_check()          # raises on a non-sm_100 device: there is no fallback to fall back to
    return torch.device('cuda:0')
