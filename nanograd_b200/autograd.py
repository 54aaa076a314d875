"""The operator ABI: ``Function`` contexts linking a result to its parents.

Same contract as the reference (nanograd/autograd.py:7-41): an op is a class with static
``forward(ctx, *raw_buffers, **kwargs)`` / ``backward(ctx, grad_raw)``; ``apply`` creates the
context, mirrors forward's keyword defaults and the call's kwargs onto it, runs forward on the
raw ``.data`` of the argument tensors and wraps the result.  Additions: the signature lookup is
cached per class (the reference calls inspect.signature on every apply), the result
records ``children`` / ``op`` so the graph visualiser has edges to draw, and the result is built
with the *caller's* Tensor class (``type(args[0])``) — the op namespace therefore works
unchanged under this package's Tensor and under the reference's own ``nanograd.tensor.Tensor``
(``nanograd_b200.integrate``), whose ``Device`` enum is a different object.
"""
import inspect

from nanograd_b200 import dist as _dist


class Function:
    def __init__(self, *tensors):
        self.parents = tensors
        self.saved_tensors = []

    def save_for_backward(self, *args):
        self.saved_tensors.extend(args)

    @staticmethod
    def forward(ctx, *args, **kwargs):
        raise NotImplementedError("Function subclasses implement forward")

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("Function subclasses implement backward")

    _defaults_cache = {}

    @classmethod
    def _forward_defaults(cls):
        cached = Function._defaults_cache.get(cls)
        if cached is None:
            params = inspect.signature(cls.forward).parameters
            cached = [(p.name, p.default) for p in params.values() if p.default is not p.empty]
            Function._defaults_cache[cls] = cached
        return cached

    def apply(self, *args, **kwargs):
        # called as op.apply(op, *tensors, **kwargs): `self` is the op class (tensor.py:487)
        tensor_cls = type(args[0])  # the host package's Tensor (this repo's or the reference's)

        ctx = self(*args)
        for name, default in self._forward_defaults():
            setattr(ctx, name, default)
        for k, v in kwargs.items():
            setattr(ctx, k, v)

        out = self.forward(ctx, *[t.data for t in args], **kwargs)
        requires_grad = any(t.requires_grad for t in args)
        if _dist._state["world"] > 1 and requires_grad:
            _dist.note_use(args)
        ret = tensor_cls(out, device=ctx.device, requires_grad=requires_grad)
        ret.op = self.__name__.lower()
        if requires_grad:
            # edges for the visualiser; a no-grad result keeps no reference to its inputs, so
            # inference chains do not pin device buffers
            ret.children = list(args)
            ret.ctx = ctx
        return ret
