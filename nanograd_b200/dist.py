"""Data-parallel training over the GPUs of one B200 node: one process per GPU.

New functionality — the reference has no multi-device support (``Device`` carries no ordinal,
nanograd/device.py:10-11; no nccl/mpi/multiprocessing anywhere).  Each rank owns GPU
``LOCAL_RANK``, so ``Device.GPU`` keeps meaning "this process's GPU" and model code is unchanged.
The batch is sharded across ranks by the caller (``shard_bounds``); parameters are replicated
(seed NumPy identically on every rank: the reference initialises weights on the host,
module.py:30-46).  Exchanges, all NCCL over NVLink/NVSwitch called from the device library:
the flat gradient bucket, summed in chunks on a communication stream while the backward pass is
still running (``GradBucket``), with the 1/world average folded into the optimizer kernel; and,
with ``sync_bn``, BatchNorm's per-channel statistics so that the step equals the reference's
single-process step on the global batch.

Rendezvous: rank 0 writes the NCCL ids to a file named after the launch, the other ranks poll it;
barriers and max-over-ranks of timings are tiny NCCL allreduces.  ``torch.distributed`` (gloo) is
used only by the CPU test modes (``use_device="emu"`` / ``False``), never on a GPU run.
"""
import os

_state = {"initialized": False, "rank": 0, "world": 1, "local_rank": 0, "nccl": False, "gloo": False, "sync_bn": False}


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


def shard_loss_weight(global_batch, world=None, rnk=None):
    """Factor a rank multiplies its shard's *mean* loss by before ``backward()`` so that the data-parallel
    step equals the single-process step on the global batch when the shards are unequal.

    The losses of this API are means over the rows they see (NLLLoss: module.py:684-686, MSELoss: :712) and the
    optimizer averages the summed gradients with 1/world; with ``n_r`` rows on rank r the global-batch gradient is
    ``sum_r (n_r / N) g_r``, i.e. each rank's loss carries ``n_r * world / N``.  Equal shards (every BASELINE
    configuration: 4096 rows over 1/2/4/8 ranks) give exactly 1.0 and need no change to the training loop."""
    world = world_size() if world is None else world
    lo, hi = shard_bounds(global_batch, world, rnk)
    return (hi - lo) * float(world) / float(global_batch)


def _rendezvous_path():
    """File through which rank 0 hands the NCCL ids to the other ranks of this launch.  The
    parent process (the torchrun agent, or whatever spawned the ranks) and MASTER_PORT identify
    the launch, so concurrent or successive runs never read each other's file."""
    import tempfile
    tag = "%s_%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", "none"), os.getppid())
    return os.path.join(os.environ.get("NANOGRAD_B200_RDZV_DIR", tempfile.gettempdir()), "nanograd_b200_rdzv_" + tag)


def _exchange_nccl_ids(rnk, timeout_s=300.0):
    """Rank 0 creates the two ncclUniqueIds and publishes them atomically (write + rename); the
    other ranks poll for the file.  No torch.distributed, no sockets of our own: NCCL's bootstrap
    does the actual rendezvous from the ids."""
    import ctypes
    import time
    from nanograd_b200 import backend as B
    path = _rendezvous_path()
    if rnk == 0:
        uid = ctypes.create_string_buffer(256)
        B.lib().ng_nccl_unique_id(uid)
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as f:
            f.write(bytes(uid.raw))
        os.replace(tmp, path)
        return bytes(uid.raw)
    deadline = time.time() + timeout_s
    while time.time() < deadline:
        try:
            with open(path, "rb") as f:
                data = f.read()
            if len(data) == 256:
                return data
        except FileNotFoundError:
            pass
        time.sleep(0.01)
    raise Exception(f"rank {rnk}: timed out waiting for rank 0's NCCL ids at {path}")


def init(use_device=True, sync_bn=False):
    """Join the data-parallel group described by RANK / WORLD_SIZE / LOCAL_RANK (what ``torchrun``
    exports; MASTER_PORT only names the launch).

    use_device=True   the rank binds GPU LOCAL_RANK and builds the NCCL communicators; every exchange
                      (gradient bucket, SyncBN statistics, barriers, max-over-ranks of timings) is NCCL,
                      called from the device library.  No torch.distributed.
    use_device="emu"  CPU tests of the same code path: the host-emulated kernel library is bound and
                      its collectives are carried by torch.distributed's gloo backend.
    use_device=False  gloo control plane only (sharding / host-side checks).
    sync_bn           BatchNorm statistics over the global batch (the reference's single-process
                      semantics at the global batch size) instead of per-shard statistics."""
    if _state["initialized"]:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rnk = int(os.environ.get("RANK", "0"))
    lrank = int(os.environ.get("LOCAL_RANK", str(rnk)))
    _state.update(rank=rnk, world=world, local_rank=lrank)
    if world > 1 and use_device is not True:
        import torch.distributed as td
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if not td.is_initialized():
            td.init_process_group(backend="gloo", rank=rnk, world_size=world)
        _state["gloo"] = True
    if use_device:
        from nanograd_b200 import backend as B
        B.init(lrank if use_device is True else 0)
        if world > 1:
            import ctypes
            if use_device is True:
                # NCCL's own log lines (the version banner under NCCL_DEBUG=VERSION/INFO) go to stderr:
                # stdout carries machine-read output (bench.py prints one JSON line)
                os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
                ids = _exchange_nccl_ids(rnk)
                B.lib().ng_nccl_init(ctypes.create_string_buffer(ids, 256), world, rnk)
                _state["nccl"] = True
                all_reduce_sum(0.0)  # first collective: every rank has read the file by now
                if rnk == 0:
                    try:
                        os.remove(_rendezvous_path())
                    except OSError:
                        pass
            else:
                _bind_emulated_collectives(B, world, rnk)
                _state["nccl"] = True
            B.lib().ng_nccl_sync_bn(1 if sync_bn else 0)
    _state["sync_bn"] = bool(sync_bn) and world > 1
    _state["initialized"] = True


_EMU_CALLBACK = []


def _bind_emulated_collectives(B, world, rnk):
    """tests only: carry the emulation library's collectives over gloo (host memory stands in for HBM)."""
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as td

    def view(ptr, count, ctype, dtype):
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctype)), shape=(int(count),)).view(dtype)

    def collective(kind, send, recv, count):
        try:
            if kind == 0:
                a = view(send, count, ctypes.c_float, np.float32)
                t = torch.from_numpy(a.copy())
                td.all_reduce(t, op=td.ReduceOp.SUM)
                a[:] = t.numpy()
            elif kind == 1:
                a = view(send, count, ctypes.c_float, np.float32)
                out = view(recv, count * world, ctypes.c_float, np.float32)
                parts = [torch.empty(int(count), dtype=torch.float32) for _ in range(world)]
                td.all_gather(parts, torch.from_numpy(a.copy()))
                out[:] = torch.cat(parts).numpy()
            else:
                a = view(send, count, ctypes.c_double, np.float64)
                t = torch.from_numpy(a.copy())
                td.all_reduce(t, op=td.ReduceOp.SUM if kind == 2 else td.ReduceOp.MAX)
                a[:] = t.numpy()
            return 0
        except Exception as exc:  # surfaced as a library error
            print("emulated collective failed:", exc, flush=True)
            return 1

    cb = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64)(collective)
    _EMU_CALLBACK.append(cb)  # keep alive
    fn = B.lib().ng_emu_set_collective
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    fn.restype = ctypes.c_int
    fn(ctypes.cast(cb, ctypes.c_void_p), world, rnk)


def sync_bn():
    return bool(_state.get("sync_bn"))


def _host_allreduce(values, op):
    """Sum (op 0) / max (op 1) of a short list of floats across ranks."""
    import ctypes
    if world_size() == 1:
        return [float(v) for v in values]
    if _state.get("nccl"):
        from nanograd_b200 import backend as B
        buf = (ctypes.c_double * len(values))(*[float(v) for v in values])
        B.lib().ng_nccl_host_allreduce(buf, len(values), op)
        return [float(v) for v in buf]
    import torch
    import torch.distributed as td
    t = torch.tensor([float(v) for v in values], dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX if op == 1 else td.ReduceOp.SUM)
    return [float(v) for v in t]


def broadcast_object(obj_list, src=0):
    """Broadcast picklable objects from ``src`` (gloo control plane only)."""
    if world_size() > 1:
        import torch.distributed as td
        td.broadcast_object_list(obj_list, src=src)
    return obj_list


def barrier():
    if world_size() > 1:
        _host_allreduce([0.0], 0)


def all_reduce_max(value):
    """max over ranks of a Python float (device-timed step durations)."""
    return _host_allreduce([value], 1)[0]


def all_reduce_sum(value):
    return _host_allreduce([value], 0)[0]


def all_reduce_host(array):
    """Sum a NumPy array across ranks on the host (checks of the DP maths in tests)."""
    if world_size() == 1:
        return array
    import numpy as np
    flat = np.asarray(array, dtype=np.float64).reshape(-1)
    out = np.empty_like(flat)
    for i in range(0, flat.size, 64):
        out[i:i + 64] = _host_allreduce(list(flat[i:i + 64]), 0)
    return out.reshape(np.shape(array))


def replica_checksum(params):
    """(sum, sum of squares) of all parameter values on this rank and whether every rank holds the same
    two numbers — a cheap cross-rank divergence check printed by bench.py at N > 1."""
    import numpy as np
    s = q = 0.0
    for p in params:
        a = p.numpy().astype(np.float64)
        s += float(a.sum())
        q += float((a * a).sum())
    mx = _host_allreduce([s, q, -s, -q], 1)
    same = (mx[0] == -mx[2]) and (mx[1] == -mx[3])
    return {"sum": s, "sum_sq": q, "identical_across_ranks": bool(same)}


class GradBucket:
    """One flat FP32 arena holding the gradients of an optimizer's parameters — the operand of the
    data-parallel allreduce.

    * Parameter gradients are **produced in place**: the device ops that compute a parameter's
      gradient (wgrad, bias / BatchNorm sums, Linear's dW) ask ``grad_slot`` for their output buffer
      and get an alias of the parameter's slot, so there is no pack copy.  A gradient that ends up
      elsewhere (a parameter used twice accumulates into a fresh buffer) is copied in at reduce time.
    * Slots are laid out in **reverse parameter order**: backward produces the last layer's gradients
      first, so the finished prefix of the arena grows contiguously and is handed to NCCL in chunks
      (``ng_nccl_allreduce_begin`` on the communication stream) while the backward pass of the earlier
      layers is still running.  ``reduce`` launches what is left and makes the compute stream wait.
    * After ``reduce`` the slots hold the cross-rank **sums** (``p.grad`` shows the sum; the optimizer
      kernel folds in 1/world).
    """

    ALIGN = 32               # floats: every slot starts on a 128-byte boundary (vectorised optimizer loads)
    # launch an allreduce once >= 128 KB of finished gradients is waiting.  (With 1 MB chunks the last three conv
    # layers of config 2 — 147 k + 74 k + 37 k floats — stayed just under the threshold and went out only after the
    # whole backward, an exposed 1 MB allreduce; now only the first layer's 1.7 k floats are left for the end.)
    CHUNK_FLOATS = int(os.environ.get("NANOGRAD_B200_DP_CHUNK_FLOATS", 1 << 15))

    def __init__(self, params):
        from nanograd_b200 import backend as B
        from nanograd_b200.nn.buffer import DeviceBuffer
        self.params = list(params)
        self.ids = [id(p) for p in self.params]
        self.numels = [int(p.data.size) for p in self.params]
        self.offsets, off = [0] * len(self.params), 0
        for i in reversed(range(len(self.params))):
            self.offsets[i] = off
            off += (self.numels[i] + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.total = off
        self.flat = DeviceBuffer((max(self.total, 1),))
        B.lib().ng_memset_zero(self.flat.ptr, self.flat.size * 4)  # padding stays zero for ever
        self._index = {pid: i for i, pid in enumerate(self.ids)}
        self._done = [False] * len(self.params)   # slot i holds its parameter's complete gradient
        self._launched = 0                        # floats [0, _launched) are already with NCCL
        for i, p in enumerate(self.params):
            p._grad_bucket = (self, i)
            p._grad_uses = 0

    def matches(self, params):
        return [id(p) for p in params] == self.ids and [int(p.data.size) for p in params] == self.numels

    def detach(self):
        for p in self.params:
            if getattr(p, "_grad_bucket", (None,))[0] is self:
                del p._grad_bucket

    def slot_ptr(self, i):
        return self.flat.ptr + 4 * self.offsets[i]

    # ---- called from the device ops' backward (nn/ops_cuda.py) --------------------------------
    def slot_for(self, i, shape):
        """Output buffer for the gradient of parameter ``i`` if this is its only contribution this
        step (first use, nothing accumulated yet), else None."""
        p = self.params[i]
        if p.grad is not None or self._done[i] or p._grad_uses != 1:
            return None
        from nanograd_b200.nn.buffer import DeviceBuffer
        if int(np_prod(shape)) != self.numels[i]:
            return None
        return DeviceBuffer.alias_at(self.flat, self.offsets[i], shape)

    def produced(self, i):
        """The kernel writing slot ``i`` has been enqueued: extend the finished prefix and launch the
        allreduce of every full chunk of it."""
        self._done[i] = True
        self._launch(final=False)

    def _launch(self, final):
        from nanograd_b200 import backend as B
        # finished prefix in arena order (= reverse parameter order)
        end = self._launched
        for i in reversed(range(len(self.params))):
            if self.offsets[i] < end:
                continue
            if not self._done[i]:
                break
            end = self.offsets[i] + (self.numels[i] + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        if end > self._launched and (final or end - self._launched >= self.CHUNK_FLOATS):
            B.lib().ng_nccl_allreduce_begin(self.flat.ptr + 4 * self._launched, end - self._launched)
            self._launched = end

    # ---- called from the optimizer step ---------------------------------------------------------
    def reduce(self, params):
        from nanograd_b200 import backend as B
        lib = B.lib()
        for i, p in enumerate(params):
            if p.grad.data.ptr != self.slot_ptr(i):
                if self.offsets[i] < self._launched:
                    raise Exception("gradient bucket: a parameter's gradient changed after its chunk was reduced")
                lib.ng_d2d(self.slot_ptr(i), p.grad.data.ptr, 4 * self.numels[i])
            self._done[i] = True
        self._launch(final=True)
        if self._launched != self.total:
            raise Exception("gradient bucket: not every slot was reduced")
        lib.ng_nccl_allreduce_end()
        self._done = [False] * len(self.params)
        self._launched = 0
        for p in self.params:
            p._grad_uses = 0
        return [self.slot_ptr(i) for i in range(len(params))]


def np_prod(shape):
    n = 1
    for s in shape:
        n *= int(s)
    return n


def grad_slot(ctx, parent_index, shape):
    """Device ops call this for the output buffer of a parameter's gradient.  Returns
    ``(buffer or None, token)``; after enqueuing the kernel that fills the buffer the op calls
    ``grad_slot_filled(token)``."""
    if _state["world"] == 1 or parent_index >= len(ctx.parents):
        return None, None
    reg = getattr(ctx.parents[parent_index], "_grad_bucket", None)
    if reg is None:
        return None, None
    bucket, i = reg
    buf = bucket.slot_for(i, shape)
    return buf, ((bucket, i) if buf is not None else None)


def grad_slot_filled(token):
    if token is not None:
        token[0].produced(token[1])


def note_use(args):
    """Function.apply reports the tensors of every op: a parameter feeding more than one op gets its
    gradient by accumulation and cannot be produced in place."""
    for t in args:
        if hasattr(t, "_grad_bucket"):
            t._grad_uses += 1


def shutdown():
    if not _state["initialized"]:
        return
    if _state["nccl"]:
        from nanograd_b200 import backend as B
        B.lib().ng_nccl_destroy()
        _state["nccl"] = False
    if _state.get("gloo"):
        import torch.distributed as td
        if td.is_initialized():
            td.destroy_process_group()
        _state["gloo"] = False
    _state.update(initialized=False, rank=0, world=1, local_rank=0, sync_bn=False)
