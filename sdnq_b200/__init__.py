"""sdnq_b200: B200-native (sm_100a) implementation of the SDNQ quantized-Linear forward."""
from .common import dtype_dict, sdnq_version  # noqa: F401

__version__ = "0.1.0"
