"""rec-attend B200: the recurrent-attention decoding hot path of renmengye/rec-attend-public
as hand-written sm_100a CUDA behind the reference's own boundaries."""
from . import config, synthetic  # noqa: F401

__all__ = ['config', 'synthetic']
