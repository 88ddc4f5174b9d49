"""pir_b200 — B200-native PIR server answer path (drop-in for OpenMined/PIR's PIRServer / PIRDatabase hot path).

The compute lives in hand-written sm_100a CUDA kernels behind the C ABI of include/pir_b200.h
(pir_b200/lib/libpirb200.so).  This package is the thin host-side mirror of the reference's interface.
"""
from .api import (INTERNAL, INVALID_ARGUMENT, CreatePIRParameters, EncryptionParameters, GaloisKeys,  # noqa: F401
                  GenerateEncryptionParams, PIRDatabase, PIRParameters, PIRServer, PIRStatusError, Request, Response,
                  StringEncoder, calculate_dimensions, ceil_log2, generate_galois_elts, log2, next_power_two)
