"""B200-native 3GPP NR LDPC engine behind the NRLDPCEncoder / NRLDPCDecoder API of
robmaunder/ldpc-3gpp-matlab.  Compute lives in libnrldpc_b200.so (hand-written sm_100a CUDA,
C ABI in include/nrldpc_b200.h); this package is the host-side mirror of the reference's
System objects plus the ctypes binding."""
from . import capi  # noqa: F401
from .capi import CudaError, NRLDPCError, UnsupportedParameters  # noqa: F401
