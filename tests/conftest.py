import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, 'oracle')
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)

REF_DATA = '/root/reference/demos/data_48k'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def ref_modules():
    """The mechanically py3-translated reference (oracle/_ref); only exists where /root/reference does."""
    ref_dir = os.path.join(ORACLE_DIR, '_ref')
    if not os.path.isdir('/root/reference/src'):
        pytest.skip('/root/reference not present (GPU box): reference cross-check skipped')
    sys.path.insert(0, ORACLE_DIR)
    import make_ref
    make_ref.build()
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import magphase as ref_mp
        import libaudio as ref_la
        import libutils as ref_lu
    return ref_mp, ref_la, ref_lu


def have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
