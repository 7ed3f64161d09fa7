"""opendpd_b200 — B200-native (sm_100a) replacement of OpenDPD's recurrent-backbone hot path.

Host side is Python (the reference is Python); all arithmetic of the hot path runs in hand-written CUDA kernels
behind the C ABI of ``libodpd.so`` (include/odpd.h).  There is NO CPU / PyTorch fallback on this path: importing
``opendpd_b200._ffi`` without the built library raises.
"""
__version__ = "0.1.0"
