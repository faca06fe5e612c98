"""B200-native LPG / GROOVE meta-training inner loop."""
import os

# GRU arithmetic of the LPG network:
#   "tc"   tcgen05 tensor cores, fp16 operands / fp32 accumulation in TMEM (production path)
#   "fp32" exact-fp32 SIMT kernels (numerical baseline)
GRU_PRECISION = os.environ.get("TOUED_GRU_PRECISION", "tc")
