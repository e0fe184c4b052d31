"""opv_cxx_demod_b200 — B200-native opv-demod receive chain (hand-written sm_100a CUDA behind a C ABI).

The package holds only what the hot path needs: csrc/ (kernels + C ABI + drop-in CLI), capi.py
(ctypes binding) and demod.py (host-side mirror of the reference's objects).  Importing it does
not load the library; the first call does, and fails loudly if libopvd.so has not been built.
"""
from .capi import OpvdError, build, BANK_CLI_PATH, CLI_PATH, LIB_PATH  # noqa: F401
from .demod import DemodBank, Frames, make_synth, stage_decode, synth_bank  # noqa: F401

__all__ = ["DemodBank", "Frames", "stage_decode", "make_synth", "synth_bank", "build", "OpvdError", "CLI_PATH",
           "BANK_CLI_PATH", "LIB_PATH"]
