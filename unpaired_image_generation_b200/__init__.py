"""B200-native CycleGAN training step (sm_100a CUDA behind a C ABI) with the module / train-step API
of the committed stand-in (oracle/cyclegan_standin.py).  No CPU fallback: importing the compute entry
points without the built libcyclegan_b200.so raises."""
import os as _os

# The step graph has up to 8 parallel branches; give the driver enough hardware work queues for them
# (must be set before the CUDA context exists; a no-op if the user already chose a value).
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .engine import LOSS_KEYS, StepEngine, describe
from .modules import Discriminator, Generator
from .parallel import GradSync
from .trainer import CycleGANTrainer

__all__ = ["Generator", "Discriminator", "CycleGANTrainer", "StepEngine", "GradSync", "LOSS_KEYS", "describe"]
