"""dicow-b200: B200-native (sm_100a) implementation of the TS-ASR-Whisper (DiCoW / SE-DiCoW) hot path.

Sub-modules
    lib      ctypes binding of libdicow_b200.so (C ABI in include/dicow_b200.h)
    ops      torch-tensor wrappers around the C ABI (pointer plumbing only)
    build    in-tree nvcc build of the library
"""
__version__ = "0.1.0"
