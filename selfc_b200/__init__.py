"""selfc_b200: B200-native (sm_100a) implementation of SelfC's 4x video-rescaling hot path behind the reference's
model API.  See DESIGN.md.  Importing the package does not load the CUDA library; the first op does, and fails
loudly if libselfc_b200.so has not been built (there is no CPU / PyTorch fallback)."""
from .global_var import GlobalVar  # noqa: F401

__all__ = ["GlobalVar", "arch", "engine", "networks", "options", "build"]
__version__ = "0.1.0"
