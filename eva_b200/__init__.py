"""eva_b200 -- B200-native (sm_100a) CKKS execution backend behind the EVA API.

The hot path (homomorphic evaluation of a compiled EVA program) runs in
hand-written CUDA kernels reached through the C-ABI in include/evab200.h
(built into eva_b200/lib/libevab200.so).  There is no CPU fallback: loading the
backend without the CUDA library or without a GPU raises.
"""
__version__ = "0.1.0"
