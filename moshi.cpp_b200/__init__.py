"""moshi.cpp_b200 — B200-native LM decode step for moshi.cpp (temporal transformer + depformer).

The product is the C-ABI shared library built from csrc/ (include/moshi_b200.h) and the C++ host
mirror of the reference API in host/.  This Python package is plumbing for tests and bench.py:
model presets, the random-init GGUF writer and a ctypes binding of the C-ABI.
"""
from . import configs, parallel, synth  # noqa: F401
