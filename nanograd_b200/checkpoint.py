"""Checkpoint / resume: the reference has no save/load of any kind.

``save(path, model, optimizer)`` writes one ``.npz`` with every parameter (in ``model.parameters()``
order, with its dotted path so that a reordered model is refused on load), BatchNorm running statistics, and the optimizer state (SGD momentum buffers, Adam/AdamW moments
and step counter); ``load`` restores them in place on whichever device the objects live on (device
tensors are overwritten through one H2D copy each, so optimizer pointer tables stay valid).
"""
import numpy as np

from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor


def _host(t):
    return np.array(t.numpy() if t.device == Device.GPU else t.data, dtype=np.float32, copy=True)


def _assign(t, value):
    value = np.ascontiguousarray(value, dtype=np.float32)
    if value.size != int(np.prod(t.shape, dtype=np.int64)):
        raise Exception(f"checkpoint tensor has shape {value.shape}, expected {t.shape}")
    # same element count, different logical shape: a 0-d parameter (Swish's beta) is stored as (1,)
    # once an optimizer step has touched it (optimizer._mirror_scalar_shapes)
    value = value.reshape(t.shape)
    if t.device == Device.GPU:
        import ctypes
        from nanograd_b200 import backend as B
        B.lib().ng_h2d(t.data.ptr, value.ctypes.data_as(ctypes.c_void_p), value.nbytes)
        B.sync()  # `value` is pageable host memory: keep it alive until the copy is over
    else:
        t.data = value


def _buffers(model):
    """Non-parameter tensors that belong to the state: BatchNorm running statistics."""
    out = []

    def visit(m, prefix):
        for name in ("running_mean", "running_var"):
            t = getattr(m, name, None)
            if isinstance(t, Tensor):
                out.append((f"{prefix}{name}", m, name))
        for key, sub in getattr(m, "_submodules", {}).items():
            visit(sub, f"{prefix}{key}.")
    visit(model, "")
    return out


_OPT_LISTS = ("momentums", "exp_avg", "exp_avg_sq")


def parameter_paths(model):
    """Dotted attribute path of every parameter, in ``model.parameters()`` order (own parameters
    first, then sub-modules: module.py:85-95)."""
    out = []

    def visit(m, prefix):
        out.extend(f"{prefix}{name}" for name in getattr(m, "_parameters", {}))
        for key, sub in getattr(m, "_submodules", {}).items():
            visit(sub, f"{prefix}{key}.")
    visit(model, "")
    return out


def save(path, model, optimizer=None):
    arrays = {f"param{i}": _host(p) for i, p in enumerate(model.parameters())}
    arrays["param_paths"] = np.array(parameter_paths(model))
    for key, mod, name in _buffers(model):
        arrays[f"buffer:{key}"] = _host(getattr(mod, name))
    if optimizer is not None:
        for lst in _OPT_LISTS:
            for i, t in enumerate(getattr(optimizer, lst, []) or []):
                arrays[f"opt:{lst}{i}"] = _host(t)
        if hasattr(optimizer, "t"):
            arrays["opt:t"] = np.array([optimizer.t], dtype=np.int64)
    np.savez(path, **arrays)


def load(path, model, optimizer=None):
    with np.load(path) as f:
        params = list(model.parameters())
        n_saved = len([k for k in f.files if k.startswith("param") and k != "param_paths"])
        if n_saved != len(params):
            raise Exception(f"checkpoint has {n_saved} parameters, the model has {len(params)}")
        if "param_paths" in f.files:
            saved, mine = [str(x) for x in f["param_paths"]], parameter_paths(model)
            if len(mine) == len(saved) and saved != mine:
                diff = next((a, b) for a, b in zip(saved, mine) if a != b)
                raise Exception(f"checkpoint parameter order differs from the model's: saved '{diff[0]}' where the "
                                f"model has '{diff[1]}'")
        for i, p in enumerate(params):
            _assign(p, f[f"param{i}"])
        for key, mod, name in _buffers(model):
            if f"buffer:{key}" in f.files:
                _assign(getattr(mod, name), f[f"buffer:{key}"])
        if optimizer is not None:
            for lst in _OPT_LISTS:
                for i, t in enumerate(getattr(optimizer, lst, []) or []):
                    if f"opt:{lst}{i}" in f.files:
                        _assign(t, f[f"opt:{lst}{i}"])
            if "opt:t" in f.files and hasattr(optimizer, "t"):
                optimizer.t = int(f["opt:t"][0])
