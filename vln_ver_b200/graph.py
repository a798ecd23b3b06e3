"""CUDA-graph capture of one pass of the hot path (training step or inference forward).

A training step of the lift+encode path is ~900 kernel launches; at small voxel grids the host cannot enqueue
them as fast as the GPU runs them (round 1: 20x20x8, batch 2 ran 6.5 ms per step, slower per panorama than
40x40x16).  The reference has the opposite problem -- host synchronisations inside SpatialCrossAttention.forward
(M/spatial_cross_attention.py:140,142: a nonzero() per camera) -- which this path already removed; what is left
is pure launch cost, and a captured graph replays it with one launch.

What makes the step capturable:
  * no host synchronisation anywhere on the path (visibility lists, sort, loss, clipping all stay on the device);
  * every libver_b200 call launches on torch's current stream with caller-owned buffers (the capture stream and
    the graph's private memory pool);
  * dropout keys: host-side keys are frozen into the graph, so the kernels add a device-side epoch word that the
    captured step bumps (ops.advance_dropout_epoch) -- fresh masks on every replay;
  * the low-precision weight copies (fused_layer.half_of) are re-made INSIDE the graph: the cache is invalidated
    right before capture, otherwise a cache hit would freeze stale copies into the graph.
"""
import torch

from . import fused_layer, ops
from ._lib import launch_count


class CapturedStep:
    """fn(**tensors) -> tensor | tuple of tensors, captured once and replayed on new input values.

    `example_inputs` fixes shapes / dtypes / device; inputs are copied into static buffers before each replay
    (stream-ordered, non-blocking), outputs are static tensors that the next replay overwrites."""

    def __init__(self, fn, example_inputs, warmup=3, ddp=False):
        self.static = {k: v.clone() for k, v in example_inputs.items()}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 11 if ddp else 1)):      # DDP wants >= 11 eager iterations before capture
                fn(**self.static)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        fused_layer.invalidate_weight_cache()
        self.graph = torch.cuda.CUDAGraph()
        n0 = launch_count()
        # NCCL's watchdog thread polls events: a thread-local capture keeps it from invalidating the capture
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local' if ddp else 'global'):
            ops.advance_dropout_epoch()
            self.out = fn(**self.static)
        self.launches_per_replay = launch_count() - n0      # libver_b200 kernels inside one replay
        torch.cuda.synchronize()

    def __call__(self, **inputs):
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out
