"""deepfir_b200 — Python binding of libdfir_b200.so, the sm_100a implementation of the Deep-FIR SISR
forward hot path (Q-RCAN trunk, meta-attention, PixelShuffle upsampler).

The binding is ctypes over the C ABI declared in ``include/dfir.h``; PyTorch only provides device
memory, streams and the custom-op registration.  There is no CPU or eager-PyTorch fallback: importing
works anywhere (so that checkpoints, configs and handlers can be built on a CPU box) but every compute
call raises ``RuntimeError`` unless it runs on a CUDA device of compute capability 10.x with the
in-tree shared library present.
"""
from ._lib import lib_path, load_library, DfirError  # noqa: F401
