"""Data-parallel training over the GPUs of one B200 node: one process per GPU.

New functionality — the reference has no multi-device support (``Device`` carries no ordinal,
nanograd/device.py:10-11; no nccl/mpi/multiprocessing anywhere).  Each rank owns GPU
``LOCAL_RANK``, so ``Device.GPU`` keeps meaning "this process's GPU" and model code is unchanged.
The batch is sharded across ranks by the caller (``shard_bounds``); parameters are replicated
(seed NumPy identically on every rank: the reference initialises weights on the host,
module.py:30-46); the only exchange is one NCCL allreduce(sum) of the flat gradient bucket per
``optimizer.step()`` over NVLink/NVSwitch, with the 1/world average folded into the optimizer
kernel (optim/optimizer.py).

Control plane (rendezvous, broadcasting the 128-byte NCCL id, barriers, max-over-ranks timing)
goes through ``torch.distributed`` with the ``gloo`` backend on the host; the data plane is NCCL
called directly from the device library (no torch tensors on the device path).
"""
import os

_state = {"initialized": False, "rank": 0, "world": 1, "local_rank": 0, "nccl": False}


def is_initialized():
    return _state["initialized"]


def rank():
    return _state["rank"]


def world_size():
    return _state["world"]


def local_rank():
    return _state["local_rank"]


def shard_bounds(global_batch, world=None, rnk=None):
    """[start, stop) rows of the global batch owned by ``rnk`` (equal shards, remainder to the
    first ranks)."""
    world = world_size() if world is None else world
    rnk = rank() if rnk is None else rnk
    base, rem = divmod(int(global_batch), int(world))
    start = rnk * base + min(rnk, rem)
    return start, start + base + (1 if rnk < rem else 0)


def init(use_device=True):
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR /
    MASTER_PORT (what ``torchrun`` exports).  With ``use_device`` the rank binds its GPU and
    builds the NCCL communicator; ``use_device=False`` sets up the host control plane only
    (CPU tests of the sharding / rendezvous logic)."""
    if _state["initialized"]:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rnk = int(os.environ.get("RANK", "0"))
    lrank = int(os.environ.get("LOCAL_RANK", str(rnk)))
    _state.update(rank=rnk, world=world, local_rank=lrank)
    if world > 1:
        import torch.distributed as td
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if not td.is_initialized():
            td.init_process_group(backend="gloo", rank=rnk, world_size=world)
    if use_device:
        from nanograd_b200 import backend as B
        B.init(lrank)
        if world > 1:
            import ctypes
            uid = ctypes.create_string_buffer(128)
            if rnk == 0:
                B.lib().ng_nccl_unique_id(uid)
            payload = [bytes(uid.raw)]
            broadcast_object(payload)
            uid = ctypes.create_string_buffer(payload[0], 128)
            B.lib().ng_nccl_init(uid, world, rnk)
            _state["nccl"] = True
    _state["initialized"] = True


def broadcast_object(obj_list, src=0):
    """Broadcast picklable objects from ``src`` (host control plane)."""
    if world_size() > 1:
        import torch.distributed as td
        td.broadcast_object_list(obj_list, src=src)
    return obj_list


def barrier():
    if world_size() > 1:
        import torch.distributed as td
        td.barrier()


def all_reduce_max(value):
    """max over ranks of a Python float (device-timed step durations)."""
    if world_size() == 1:
        return float(value)
    import torch
    import torch.distributed as td
    t = torch.tensor([float(value)], dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def all_reduce_sum(value):
    if world_size() == 1:
        return float(value)
    import torch
    import torch.distributed as td
    t = torch.tensor([float(value)], dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return float(t.item())


def all_reduce_host(array):
    """Sum a NumPy array across ranks on the host (used by the CPU tests of the DP maths)."""
    if world_size() == 1:
        return array
    import torch
    import torch.distributed as td
    t = torch.from_numpy(array.copy())
    td.all_reduce(t, op=td.ReduceOp.SUM)
    return t.numpy()


class GradBucket:
    """One flat FP32 arena holding the gradients of an optimizer's parameters, the operand of the
    data-parallel allreduce.  ``reduce(params)`` returns the per-parameter device pointers of
    the cross-rank *sums* (the optimizer kernel folds in 1/world)."""

    ALIGN = 32  # floats: every slot starts on a 128-byte boundary (vectorised optimizer loads)

    def __init__(self, params):
        from nanograd_b200.nn.buffer import DeviceBuffer
        self.ids = [id(p) for p in params]
        self.numels = [int(p.data.size) for p in params]
        self.offsets, off = [], 0
        for n in self.numels:
            self.offsets.append(off)
            off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.total = off
        self.flat = DeviceBuffer((max(self.total, 1),))
        from nanograd_b200 import backend as B
        B.lib().ng_memset_zero(self.flat.ptr, self.flat.size * 4)  # padding stays zero for ever

    def matches(self, params):
        return [id(p) for p in params] == self.ids and [int(p.data.size) for p in params] == self.numels

    def slot_ptr(self, i):
        return self.flat.ptr + 4 * self.offsets[i]

    def reduce(self, params):
        from nanograd_b200 import backend as B
        lib = B.lib()
        for i, p in enumerate(params):
            if p.grad.data.ptr != self.slot_ptr(i):
                lib.ng_d2d(self.slot_ptr(i), p.grad.data.ptr, 4 * self.numels[i])
        lib.ng_nccl_allreduce_sum(self.flat.ptr, self.total)
        return [self.slot_ptr(i) for i in range(len(params))]


def shutdown():
    if not _state["initialized"]:
        return
    if _state["nccl"]:
        from nanograd_b200 import backend as B
        B.lib().ng_nccl_destroy()
        _state["nccl"] = False
    if _state["world"] > 1:
        import torch.distributed as td
        if td.is_initialized():
            td.destroy_process_group()
    _state.update(initialized=False, rank=0, world=1, local_rank=0)
