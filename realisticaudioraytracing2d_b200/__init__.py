"""B200-native 2-D acoustic ray tracing + impulse-response convolution (hot path only).

Compute lives in csrc/ (hand-written sm_100a CUDA behind the C-ABI of include/rar2d.h); this
package holds the ctypes binding and the host-side mirror of the reference's components.
"""
__version__ = "0.1.0"
