"""Input pipeline and train / evaluate loops.

Mirrors the reference's loops (tests/helpers.py:287-335: random mini-batches with
``np.random.randint``, ``model(Xb)`` -> ``zero_grad`` -> ``criterion()(out, Yb)`` -> ``backward`` ->
``step``, argmax accuracy) and its MNIST idx-gz reader (tests/test_mnist.py:20-38), with the two
per-step synchronisations of the reference removed on ``Device.GPU``:

* batches are gathered into pinned staging memory and copied to double-buffered device tensors on a
  copy stream while the previous step computes (``ng_h2d_prefetch`` / ``ng_prefetch_wait``);
* the accuracy is computed on the device (``ng_argmax_rows`` + ``ng_count_equal``); loss and accuracy
  are read back every ``log_every`` steps instead of every step.
"""
import gzip

import numpy as np

from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor


def load_mnist_idx(paths):
    """(X_train, Y_train, X_test, Y_test) from the four idx .gz files, scaled to [0, 1] like
    tests/test_mnist.py:20-38 (16-byte image header, 8-byte label header)."""
    arrays = []
    for path in paths:
        with open(path, "rb") as f:
            arrays.append(np.frombuffer(gzip.decompress(f.read()), dtype=np.uint8))
    x_train, y_train, x_test, y_test = arrays
    x_train = x_train[0x10:].reshape((-1, 784)).astype(np.float32) / np.float32(255.0)
    x_test = x_test[0x10:].reshape((-1, 784)).astype(np.float32) / np.float32(255.0)
    return x_train, y_train[8:], x_test, y_test[8:]


class BatchLoader:
    """Yields ``steps`` mini-batches (Tensor x, Tensor y) on ``device``.

    ``indices`` is either None (the reference's sampling: ``np.random.randint(0, N, batch_size)`` per
    step, drawn from NumPy's global generator so a seeded run sees the reference's batches) or a
    callable ``step -> index array`` (sequential evaluation batches).
    """

    def __init__(self, X, Y, batch_size, steps, device=Device.CPU, indices=None):
        self.X = np.ascontiguousarray(X, dtype=np.float32)
        self.Y = None if Y is None else np.ascontiguousarray(Y, dtype=np.float32)
        self.batch_size, self.steps, self.device, self.indices = int(batch_size), int(steps), device, indices

    def _sample(self, step):
        if self.indices is not None:
            return np.asarray(self.indices(step))
        return np.random.randint(0, self.X.shape[0], size=(self.batch_size,))

    def __iter__(self):
        # index arrays are drawn up front, in step order, on the caller's thread: the sequence of
        # np.random.randint calls is the reference's
        plan = [np.ascontiguousarray(self._sample(step), dtype=np.int64) for step in range(self.steps)]
        if self.device != Device.GPU:
            for idx in plan:
                yield Tensor(self.X[idx]), (None if self.Y is None else Tensor(self.Y[idx]))
            return
        # Device.GPU: a host thread owned by the device library (ng_loader_*) gathers batch t+1 into
        # pinned memory and copies it on the copy stream into the other of two device slots while
        # step t computes; next() is one C call.
        import ctypes
        from nanograd_b200 import backend as B
        from nanograd_b200.nn.buffer import DeviceBuffer
        lib = B.lib()
        xs = (self.batch_size,) + self.X.shape[1:]
        x_dev = [DeviceBuffer(xs), DeviceBuffer(xs)]
        y_dev = [DeviceBuffer((self.batch_size,)), DeviceBuffer((self.batch_size,))] if self.Y is not None else [None, None]
        flat = np.concatenate(plan) if plan else np.zeros(0, dtype=np.int64)
        offsets = np.zeros(self.steps + 1, dtype=np.int64)
        np.cumsum([len(p) for p in plan], out=offsets[1:])
        handle = ctypes.c_uint64(0)
        row_bytes = int(np.prod(self.X.shape[1:])) * 4
        lib.ng_loader_create(ctypes.byref(handle), self.X.ctypes.data_as(ctypes.c_void_p),
                             self.Y.ctypes.data_as(ctypes.c_void_p) if self.Y is not None else None,
                             self.X.shape[0], row_bytes, self.batch_size, self.steps,
                             flat.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                             offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                             x_dev[0].ptr, x_dev[1].ptr,
                             y_dev[0].ptr if self.Y is not None else 0, y_dev[1].ptr if self.Y is not None else 0)
        slot, rows = ctypes.c_int(0), ctypes.c_int64(0)
        try:
            for _ in range(self.steps):
                lib.ng_loader_next(handle.value, ctypes.byref(slot), ctypes.byref(rows))
                k, n = slot.value, rows.value
                if n == self.batch_size:
                    # zero-copy: the Tensors wrap the slot's device buffers (valid until the next-but-one batch)
                    xb = Tensor(x_dev[k], device=Device.GPU)
                    yb = None if self.Y is None else Tensor(y_dev[k], device=Device.GPU)
                else:
                    # ragged last evaluation batch: exactly-sized device tensors (device-to-device copy)
                    xb = Tensor(DeviceBuffer((n,) + self.X.shape[1:]), device=Device.GPU)
                    lib.ng_d2d(xb.data.ptr, x_dev[k].ptr, n * row_bytes)
                    yb = None
                    if self.Y is not None:
                        yb = Tensor(DeviceBuffer((n,)), device=Device.GPU)
                        lib.ng_d2d(yb.data.ptr, y_dev[k].ptr, n * 4)
                yield xb, yb
        finally:
            lib.ng_loader_destroy(handle.value)


def batch_accuracy(out, yb):
    """Fraction of rows of ``out`` whose argmax equals the label, as a Python float.  On Device.GPU
    the argmax and the comparison run on the device and 4 bytes are read back."""
    if out.device == Device.GPU:
        from nanograd_b200 import backend as B
        from nanograd_b200.nn.buffer import DeviceBuffer
        rows, cols = out.shape[0], int(np.prod(out.shape[1:]))
        pred, cnt = DeviceBuffer((rows,)), DeviceBuffer((1,))
        B.lib().ng_argmax_rows(pred.ptr, out.data.ptr, rows, cols)
        B.lib().ng_count_equal(cnt.ptr, pred.ptr, yb.data.ptr, rows)
        return float(cnt.numpy()[0]) / rows
    cat = np.argmax(np.asarray(out.data), axis=-1)
    return float((cat == np.asarray(yb.data)).mean())


def train(model, X, y, optimizer, steps=1000, batch_size=128, criterion=None, device=Device.CPU, log_every=50,
          callback=None):
    """tests/helpers.py:287-317.  Returns (losses, accuracies) sampled every ``log_every`` steps."""
    import nanograd_b200.nn.module as nn
    criterion = criterion or nn.NLLLoss
    model.train()
    if device == Device.GPU:
        model.gpu()
    losses, accuracies = [], []
    loader = BatchLoader(X, y, batch_size, steps, device=device)
    for step, (xb, yb) in enumerate(loader):
        out = model(xb)
        optimizer.zero_grad()
        loss = criterion()(out, yb)
        loss.backward()
        optimizer.step()
        if log_every and (step % log_every == 0 or step == steps - 1):
            losses.append(float(np.asarray(loss.numpy() if hasattr(loss, "numpy") else loss.data).reshape(-1)[0]))
            accuracies.append(batch_accuracy(out, yb))
            if callback:
                callback(step, losses[-1], accuracies[-1])
    return losses, accuracies


def evaluate(model, X, y, batch_size=128, device=Device.CPU):
    """tests/helpers.py:319-335: accuracy over the whole set, sequential batches."""
    model.eval()
    if device == Device.GPU:
        model.gpu()
    n = X.shape[0]
    steps = (n - 1) // batch_size + 1
    loader = BatchLoader(X, y, batch_size, steps, device=device,
                         indices=lambda i: np.arange(batch_size * i, min(batch_size * (i + 1), n)))
    correct = 0.0
    for xb, yb in loader:
        correct += batch_accuracy(model(xb), yb) * xb.shape[0]
    return correct / n
