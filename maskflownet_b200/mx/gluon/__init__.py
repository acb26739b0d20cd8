"""mxnet.gluon stand-in: only `nn` (blocks) is provided -- training (Trainer, autograd) belongs to the caller's torch code."""
from . import nn  # noqa: F401
from .nn import Block, HybridBlock, Parameter  # noqa: F401
