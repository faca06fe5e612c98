"""B200-native LPG / GROOVE meta-training inner loop."""
import os

# GRU arithmetic of the LPG network:
#   "tc"   tcgen05 tensor cores, fp16 operands / fp32 accumulation in TMEM (production path)
#   "fp32" exact-fp32 SIMT kernels (numerical baseline)
GRU_PRECISION = os.environ.get("TOUED_GRU_PRECISION", "tc")
# per-candidate LPG forward of the ES path (meta/es.py): "tc" = tensor-core kernel with one parameter set and one set of
# pass images per CTA (default, 8x faster), "fp32" = exact SIMT kernel (tight parity tests)
ES_PRECISION = os.environ.get("TOUED_ES_PRECISION", "tc")

# Number of CUDA streams on which independent mini-batches of agents run concurrently inside one meta-step
# (only used when num_mini_batches > 1; results are independent of it).
NUM_STREAMS = int(os.environ.get("TOUED_NUM_STREAMS", "4"))
# False: every kernel of a chunk runs on the chunk's own stream (no side streams for the token sort, the agent adjoint,
# the embedding gradient and eval_agent) -- used by bench.py's per-kernel timing pass so that no launch has a neighbour
SIDE_STREAMS = os.environ.get("TOUED_SIDE_STREAMS", "1") != "0"
# True (default): make_lpg_train_step returns a step that is captured into a CUDA graph on its second call and replayed
# afterwards (meta/graph.py; the states it returns alias static buffers).  "0": every call enqueues its launches eagerly.
CUDA_GRAPH = os.environ.get("TOUED_CUDA_GRAPH", "1") != "0"
# When a list: the meta-gradient step appends (label, chunk, timing-enabled CUDA event) at its phase boundaries
# (tools/phase_timeline.py); None in production.
PHASE_EVENTS = None
