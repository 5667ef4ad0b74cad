"""p25rx_b200 -- B200-native (sm_100a) P25 Phase 1 baseband hot path.

CUDA kernels + C ABI in csrc/ (libp25cu.so); this package mirrors the reference's
DemodTask / MessageReceiver / ReplayReceiver interfaces over that ABI.
"""
from ._lib import EVENT_DTYPE, FMT_CF32_IQ, FMT_U8_IQ, P25Error, build  # noqa: F401
from .pipeline import (Context, DemodTask, MessageReceiver, ReplayReceiver, EVENT_NAMES, STATS_FAMILIES,  # noqa: F401
                       EV_ERROR, EV_NID, EV_VOICE_HEADER, EV_LINK_CONTROL, EV_CRYPTO_CONTROL, EV_LSD,
                       EV_VOICE_FRAME, EV_TSBK, EV_VOICE_TERM)
