"""maskflownet_b200 -- Blackwell-native (sm_100a) implementation of the MaskFlownet hot path:
the local correlation cost volume and the flow-guided deformable feature warp x occlusion mask that feeds it.

  maskflownet_b200.ops      tensor-level operators backed by libmaskflow_b200.so (C ABI: include/maskflow_b200.h)
  maskflownet_b200.mx       torch-backed shim of the MXNet `F` / gluon `nn` namespaces the reference's
                            network/MaskFlownet.py and network/layer.py are written against
  maskflownet_b200.network  host-side mirror of the reference's model graph that calls the fused operators
"""
from . import _lib  # noqa: F401
from ._lib import MaskflowError  # noqa: F401

__version__ = "0.1.0"
