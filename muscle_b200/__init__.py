"""muscle_b200 -- B200-native pair engine for MUSCLE5's MPCFlat stage.

engine.Engine    : ctypes binding of libmuscle_b200.so (include/muscle_b200.h)
mpcflat.MPCFlat  : host-side mirror of the reference's MPCFlat call surface (tests, bench)
dist             : multi-GPU plumbing (pair sharding, NCCL all-gather-v of the sparse store)
synth            : deterministic synthetic protein families (BASELINE.json configs C1..C5)
The CUDA sources live in muscle_b200/csrc (build: make -C muscle_b200/csrc, sm_100a only).
"""
__version__ = "0.1.0"
