import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def test_weights(seed: int = 0, gain: float = 1.0):
    """Weights for the full-size parity tests: the real SelfC-large checkpoint when one is present on this machine
    ($SELFC_CKPT or pretrained_models/selfc_large_pretrain.pth, base_model.py:87-107), the oracle's seeded weights otherwise."""
    from oracle import selfc_oracle as so
    from selfc_b200.synthetic import checkpoint_path, load_checkpoint
    path = checkpoint_path()
    if path is not None:
        sd = load_checkpoint(path)
        missing = [k for k in so.param_shapes() if k not in sd]
        if missing:
            raise RuntimeError(f"checkpoint {path} lacks {len(missing)} SelfCInvNet tensors, e.g. {missing[:3]}")
        return sd, f"checkpoint {path}"
    return so.make_state_dict(seed, gain), f"seeded random (seed {seed}, gain {gain})"


test_weights.__test__ = False
