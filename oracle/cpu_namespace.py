"""TEST INFRASTRUCTURE — the oracle as a ``Device.CPU`` operator namespace.

Wraps oracle/np_ops.py in ``Function`` classes with the reference's 21 names and CPU-side
signatures (nanograd/nn/ops_cpu.py) so that tests and the CPU-baseline leg of bench.py can run
the *same model code* on both devices:

    from nanograd_b200.tensor import register_ops
    from oracle import cpu_namespace
    register_ops(cpu_namespace, device=Device.CPU)     # tests / bench only

With no fused ops in this namespace, the device-agnostic layers compose primitives exactly as
the reference does (tensor.py:327-331,370-406; module.py:356-391,476,686; optimizer.py), which is
what makes ``Tensor`` + this namespace a faithful stand-in for the reference's NumPy path.
"""
import numpy as np

from nanograd_b200.autograd import Function
from oracle import np_ops as K


class OneHot(Function):
    @staticmethod
    def forward(ctx, input, num_classes):
        return K.onehot_forward(input, num_classes)

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError


class Unsqueeze(Function):
    @staticmethod
    def forward(ctx, input, axis):
        ctx.save_for_backward(axis)
        return np.expand_dims(input, axis)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.squeeze(ctx.saved_tensors[0])


class Squeeze(Function):
    @staticmethod
    def forward(ctx, input, axis):
        ctx.save_for_backward(axis)
        return np.squeeze(input, axis)

    @staticmethod
    def backward(ctx, grad_output):
        return np.expand_dims(grad_output, ctx.saved_tensors[0])


class Slice(Function):
    @staticmethod
    def forward(ctx, input, indices):
        ctx.save_for_backward(input.shape, indices)
        return K.slice_forward(input, indices)

    @staticmethod
    def backward(ctx, grad_output):
        shape, indices = ctx.saved_tensors
        return K.slice_backward(grad_output, shape, indices)


class Transpose(Function):
    @staticmethod
    def forward(ctx, input):
        return input.T

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.T


class Reshape(Function):
    @staticmethod
    def forward(ctx, input, shape):
        ctx.save_for_backward(input.shape)
        return input.reshape(shape)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.reshape(ctx.saved_tensors[0])


class Max(Function):
    @staticmethod
    def forward(ctx, input, axis):
        out, keep, axis = K.extreme_forward(input, axis, "max")
        ctx.save_for_backward(input, axis, keep)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, axis, keep = ctx.saved_tensors
        return K.extreme_backward(grad_output, input, axis, keep)


class Min(Function):
    @staticmethod
    def forward(ctx, input, axis):
        out, keep, axis = K.extreme_forward(input, axis, "min")
        ctx.save_for_backward(input, axis, keep)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, axis, keep = ctx.saved_tensors
        return K.extreme_backward(grad_output, input, axis, keep)


class Sum(Function):
    @staticmethod
    def forward(ctx, input, axis=None):
        ctx.save_for_backward(input, axis)
        return K.sum_forward(input, axis)

    @staticmethod
    def backward(ctx, grad_output):
        input, axis = ctx.saved_tensors
        return K.sum_backward(grad_output, input, axis)


class Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a.shape, b.shape)
        return a + b

    @staticmethod
    def backward(ctx, grad_output):
        return K.add_backward(grad_output, *ctx.saved_tensors)


class Mul(Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return a * b

    @staticmethod
    def backward(ctx, grad_output):
        return K.mul_backward(grad_output, *ctx.saved_tensors)


class MatMul(Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return a @ b

    @staticmethod
    def backward(ctx, grad_output):
        return K.matmul_backward(grad_output, *ctx.saved_tensors)


class Log(Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return np.log(input)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output / ctx.saved_tensors[0]


class Exp(Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return np.exp(input)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output * np.exp(ctx.saved_tensors[0])


class Neg(Function):
    @staticmethod
    def forward(ctx, input):
        return -input

    @staticmethod
    def backward(ctx, grad_output):
        return -grad_output


class Pow(Function):
    @staticmethod
    def forward(ctx, input, power):
        ctx.save_for_backward(input, power)
        return input ** power

    @staticmethod
    def backward(ctx, grad_output):
        return K.pow_backward(grad_output, *ctx.saved_tensors)


class ReLU(Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return np.maximum(input, 0)

    @staticmethod
    def backward(ctx, grad_output):
        return K.relu_backward(grad_output, ctx.saved_tensors[0])


class Sigmoid(Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return K.sigmoid(input)

    @staticmethod
    def backward(ctx, grad_output):
        return K.sigmoid_backward(grad_output, ctx.saved_tensors[0])


class Tanh(Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return np.tanh(input)

    @staticmethod
    def backward(ctx, grad_output):
        return K.tanh_backward(grad_output, ctx.saved_tensors[0])


def _stride_value(stride):
    return stride if isinstance(stride, int) else int(stride[0])


class Conv1d(Function):
    @staticmethod
    def forward(ctx, input, weight, stride):
        stride = _stride_value(stride)
        y, cols = K.conv1d_forward(input, weight, stride)
        ctx.save_for_backward(input, weight, stride, cols)
        return y

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, stride, cols = ctx.saved_tensors
        return K.conv1d_backward(grad_output, input, weight, stride, cols)


class Conv2d(Function):
    @staticmethod
    def forward(ctx, input, weight, stride):
        stride = _stride_value(stride)
        y, cols = K.conv2d_forward(input, weight, stride)
        ctx.save_for_backward(input, weight, stride, cols)
        return y

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, stride, cols = ctx.saved_tensors
        return K.conv2d_backward(grad_output, input, weight, stride, cols)
