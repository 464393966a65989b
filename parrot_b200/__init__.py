"""parrot_b200: Blackwell-native (sm_100a) implementation of the attention-RNN acoustic-feature
hot path of sotelo/parrot (model.py:Parrot), behind the reference's own API.  See DESIGN.md."""
from .model import Parrot  # noqa: F401

__all__ = ['Parrot']
