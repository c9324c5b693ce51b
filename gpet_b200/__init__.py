"""gpet_b200 -- B200-native (sm_100a) implementation of gPET's Monte-Carlo hot path behind a C ABI.

`gpet_b200.api` is the ctypes face of libgpet_b200.so; `gpet_b200.refio` reads/writes the reference's file formats.
"""
from .api import Context, GpetError, lib, EVENT_DTYPE, COINC_DTYPE, HIT_DTYPE, PHOTON_DTYPE, PANEL_DTYPE  # noqa: F401
