"""Captured training steps: record one step on the compute stream, replay it with one launch.

New functionality for BASELINE config 0.  The reference's loop (README.md:94-125,
tests/helpers.py:287-317) dispatches every primitive of a step from Python; on a B200 the README
MNIST CNN at batch 128 is ~30 kernels of a few microseconds each, so the step time is Python and
launch overhead (measured 0.37-0.42 ms per step for ~40 us of device work).  ``GraphedStep`` runs the
user's step function once under CUDA stream capture and afterwards replays the resulting graph:

    def step(xb, yb):                       # any code over device Tensors, no host reads
        out = model(xb)
        optimizer.zero_grad()
        loss = criterion(out, yb)
        loss.backward()
        optimizer.step()
        return loss

    fast = GraphedStep(step, X[:128], Y[:128])       # example host batches fix the shapes
    for xb, yb in batches:
        loss = fast(xb, yb)                            # H2D of the batch + one cudaGraphLaunch
    print(loss.numpy())                                # results live in fixed device buffers

Rules of a captured step: the same shapes every call; no ``.numpy()`` / ``.cpu()`` inside; host scalars
are baked in, so optimizers whose scalars change per step (Adam's bias correction) are refused; the
model's parameters, optimizer state and BatchNorm running statistics are updated in place exactly as
in the eager step (tests compare the two bit for bit).  Memory allocated during capture belongs to
the graph (csrc/runtime.cu: private pool) and is released by ``close()``.
"""
import ctypes

import numpy as np

from nanograd_b200 import backend as B
from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor


class GraphedStep:
    def __init__(self, step_fn, *example_inputs, warmup=2):
        from nanograd_b200 import dist
        if dist.is_initialized() and dist.world_size() > 1:
            raise Exception("GraphedStep: data-parallel steps are not captured (NCCL runs on its own stream)")
        B.init()
        self._fn = step_fn
        self._host = [np.ascontiguousarray(a, dtype=np.float32) for a in example_inputs]
        # two pinned staging buffers per input: the copy of call i may still be in flight when call i+1 fills
        self._pinned = [[B.pinned_empty(a.shape), B.pinned_empty(a.shape)] for a in self._host]
        self._events = [None, None]
        self._turn = 0
        self.inputs = [Tensor(a, device=Device.GPU) for a in self._host]
        for _ in range(max(int(warmup), 1)):       # lazy initialisation, pool warm-up, optimizer state placement
            step_fn(*self.inputs)
        B.sync()
        lib = B.lib()
        lib.ng_graph_begin()
        try:
            out = step_fn(*self.inputs)
        except Exception:
            h, n, b = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
            try:
                lib.ng_graph_end(ctypes.byref(h), ctypes.byref(n), ctypes.byref(b))
                lib.ng_graph_destroy(h.value)
            except Exception:
                pass
            raise
        h, n, b = ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        lib.ng_graph_end(ctypes.byref(h), ctypes.byref(n), ctypes.byref(b))
        self.handle, self.kernel_nodes, self.private_bytes = h.value, int(n.value), int(b.value)
        self.outputs = out          # static device tensors, rewritten by every replay
        lib.ng_graph_launch(self.handle)   # the capture itself executed nothing: run the step it recorded
        B.sync()

    def __call__(self, *arrays):
        if len(arrays) != len(self.inputs):
            raise Exception(f"GraphedStep expects {len(self.inputs)} inputs, got {len(arrays)}")
        lib = B.lib()
        turn = self._turn
        ev = self._events[turn]
        if ev is not None:
            ev.synchronize()        # the staging buffers of this turn are free again
        for t, stage, a in zip(self.inputs, self._pinned, arrays):
            a = np.asarray(a)
            if tuple(a.shape) != tuple(t.shape):
                raise Exception(f"GraphedStep was captured for input shape {tuple(t.shape)}, got {tuple(a.shape)}")
            np.copyto(stage[turn], a, casting="same_kind")
            lib.ng_h2d(t.data.ptr, stage[turn].ctypes.data_as(ctypes.c_void_p), stage[turn].nbytes)
        self._events[turn] = (ev or B.Event()).record()
        self._turn = 1 - turn
        lib.ng_graph_launch(self.handle)
        return self.outputs

    def close(self):
        if getattr(self, "handle", 0):
            B.lib().ng_graph_destroy(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
