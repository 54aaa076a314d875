"""Parity cases written once against nanograd's public API and run on three back ends:

  * the real reference imported from /root/reference      -> tests/golden/*.npz (make_golden.py)
  * this repo's Tensor + the NumPy oracle on Device.CPU   -> pinned against the golden vectors
  * this repo's Tensor + the sm_100a library on Device.GPU-> compared with golden + oracle

Each case takes a framework handle ``fw`` (Tensor class, nn module, optim module, device, and a
``numpy(t)`` accessor) and returns ``{name: ndarray}``.  Cases use only API that exists in the
reference at HEAD (e.g. conv2d always gets an explicit stride, slices are full rank).
Inputs come from ``np.random.RandomState(seed)`` so all three runs see identical data.
"""
import numpy as np


class Framework:
    def __init__(self, Tensor, nn, optim, device, to_numpy):
        self.Tensor, self.nn, self.optim, self.device, self.numpy = Tensor, nn, optim, device, to_numpy

    def tensor(self, array, requires_grad=False):
        t = self.Tensor(np.array(array, copy=True), requires_grad=requires_grad)
        return self.place(t)

    def place(self, obj):
        """Move a Tensor or Module to the framework's device (Device.CPU objects stay put)."""
        if self.device.name == "GPU":
            obj.gpu()
        return obj


def _grad(fw, t):
    return fw.numpy(t.grad)


def _weighted_sum(fw, out, rng):
    """loss = sum(out * m) with a fixed random m: gives every op a non-uniform upstream gradient."""
    m = fw.tensor(rng.normal(0, 1, out.shape).astype(np.float32))
    return (out * m).sum()


def _tie_rich(rng, shape):
    """Post-ReLU-like data quantised to halves: exact zeros and repeated maxima are common, which
    pins Max.backward's tie splitting (ops_cpu.py:121-127)."""
    x = np.maximum(rng.normal(0, 1, shape), 0)
    return (np.round(x * 2) / 2).astype(np.float32)


# ----------------------------------------------------------------------------- op cases
def case_conv2d(fw, k, s, seed=0, loc=0.0):
    rng = np.random.RandomState(seed + 10 * k + s)
    x = fw.tensor(rng.normal(loc, 1, (4, 3, 12, 12)).astype(np.float32), True)
    w = fw.tensor(rng.normal(loc, 1, (5, 3, k, k)).astype(np.float32), True)
    y = x.conv2d(w, s)
    _weighted_sum(fw, y, rng).backward()
    return {"y": fw.numpy(y), "dx": _grad(fw, x), "dw": _grad(fw, w)}


def case_conv2d_wide(fw, c=32, k=64, hw=16, n=4, pad=0, seed=0):
    """Conv2d shapes the tensor-core (TF32) path accepts: channels multiple of 32, stride 1.
    pad > 0 goes through nn.Conv2d (pad2d + conv2d + bias, module.py:437-476)."""
    rng = np.random.RandomState(seed + c + k + hw)
    np.random.seed(seed + 1)
    x = fw.tensor(rng.normal(0, 1, (n, c, hw, hw)).astype(np.float32), True)
    if pad:
        conv = fw.nn.Conv2d(c, k, 3, 1, pad)
        fw.place(conv)
        y = conv(x)
        w = conv.weight
    else:
        w = fw.tensor((rng.normal(0, 1, (k, c, 3, 3)) / np.sqrt(9 * c)).astype(np.float32), True)
        y = x.conv2d(w, 1)
    _weighted_sum(fw, y, rng).backward()
    return {"y": fw.numpy(y), "dx": _grad(fw, x), "dw": _grad(fw, w)}


def case_conv1d(fw, k, s, seed=0, loc=0.0):
    rng = np.random.RandomState(seed + 100 + 10 * k + s)
    x = fw.tensor(rng.normal(loc, 1, (5, 3, 40)).astype(np.float32), True)
    w = fw.tensor(rng.normal(loc, 1, (6, 3, k)).astype(np.float32), True)
    y = x.conv1d(w, s)
    _weighted_sum(fw, y, rng).backward()
    return {"y": fw.numpy(y), "dx": _grad(fw, x), "dw": _grad(fw, w)}


def case_matmul(fw, seed=1):
    rng = np.random.RandomState(seed)
    a = fw.tensor(rng.normal(0, 1, (37, 53)).astype(np.float32), True)
    b = fw.tensor(rng.normal(0, 1, (53, 29)).astype(np.float32), True)
    y = a @ b
    _weighted_sum(fw, y, rng).backward()
    return {"y": fw.numpy(y), "da": _grad(fw, a), "db": _grad(fw, b)}


def case_pool2d(fw, kind, p, seed=2):
    rng = np.random.RandomState(seed + p)
    x = fw.tensor(_tie_rich(rng, (3, 4, 11, 13)), True)
    y = x.max_pool2d((p, p)) if kind == "max" else x.avg_pool2d((p, p))
    _weighted_sum(fw, y, rng).backward()
    return {"y": fw.numpy(y), "dx": _grad(fw, x)}


def case_pool1d(fw, kind, p, seed=3):
    rng = np.random.RandomState(seed + p)
    x = fw.tensor(_tie_rich(rng, (3, 4, 29)), True)
    y = x.max_pool1d(p) if kind == "max" else x.avg_pool1d(p)
    _weighted_sum(fw, y, rng).backward()
    return {"y": fw.numpy(y), "dx": _grad(fw, x)}


def case_logsoftmax_nll(fw, seed=4):
    rng = np.random.RandomState(seed)
    x = fw.tensor(rng.normal(0, 3, (16, 10)).astype(np.float32), True)
    labels = fw.tensor(rng.randint(0, 10, (16,)).astype(np.float32))
    logp = x.log_softmax()
    loss = fw.nn.NLLLoss()(logp, labels)
    loss.backward()
    onehot = labels.onehot(num_classes=10)
    return {"logp": fw.numpy(logp), "loss": fw.numpy(loss), "dx": _grad(fw, x), "onehot": fw.numpy(onehot)}


def case_elementwise(fw, seed=5):
    rng = np.random.RandomState(seed)
    a = fw.tensor(rng.normal(3, 1, (6, 5, 4)).astype(np.float32), True)
    b = fw.tensor(rng.normal(3, 1, (6, 1, 4)).astype(np.float32), True)
    out = {}
    y = (a + b) * b - a / b + (2.5 - a) * 0.5 + a ** 3 + b ** -2 + a ** 0.3
    _weighted_sum(fw, y, rng).backward()
    out.update(y=fw.numpy(y), da=_grad(fw, a), db=_grad(fw, b))
    c = fw.tensor(rng.normal(0, 1, (7, 9)).astype(np.float32), True)
    # exact zeros pin relu'(0) = 1 (ops_cpu.py:263)
    cz = fw.numpy(c).copy()
    cz[::2, ::3] = 0.0
    c = fw.tensor(cz, True)
    z = c.relu() + c.tanh() * c.sigmoid() + c.exp() + (c * c + 1.0).log() + (c * c + 1.0).sqrt() + (-c)
    _weighted_sum(fw, z, rng).backward()
    out.update(z=fw.numpy(z), dc=_grad(fw, c))
    return out


def case_reductions(fw, seed=6):
    rng = np.random.RandomState(seed)
    x = fw.tensor(_tie_rich(rng, (5, 6, 7, 4)), True)
    s_all = x.sum()
    s_1 = x.sum(axis=1)
    s_12 = x.sum(axis=(1, 2))
    mx = x.max(axis=3).max(axis=0)
    mn = x.min(axis=(1, 3))
    me = x.mean(axis=0)
    loss = (s_all * 0.01 + _weighted_sum(fw, s_1, rng) + _weighted_sum(fw, s_12, rng) + _weighted_sum(fw, mx, rng)
            + _weighted_sum(fw, mn, rng) + _weighted_sum(fw, me, rng))
    loss.backward()
    return {"s_all": fw.numpy(s_all), "s_1": fw.numpy(s_1), "s_12": fw.numpy(s_12), "mx": fw.numpy(mx),
            "mn": fw.numpy(mn), "me": fw.numpy(me), "dx": _grad(fw, x)}


def case_shapes(fw, seed=7):
    rng = np.random.RandomState(seed)
    x = fw.tensor(rng.normal(0, 1, (6, 5, 8, 7)).astype(np.float32), True)
    sl = x[1:4, 2:, :5, 0:-1]
    pad = x.pad2d((1, 2, 3, 4))
    tr = x.reshape(shape=[30, 56]).T
    sq = x.reshape(shape=[6, 5, 56, 1]).squeeze(axis=3).unsqueeze(axis=1)
    loss = _weighted_sum(fw, sl, rng) + _weighted_sum(fw, pad, rng) + _weighted_sum(fw, tr, rng) + \
        _weighted_sum(fw, sq, rng)
    loss.backward()
    x1 = fw.tensor(rng.normal(0, 1, (4, 3, 20)).astype(np.float32), True)
    p1 = x1.pad1d((2, 3))
    _weighted_sum(fw, p1, rng).backward()
    return {"sl": fw.numpy(sl), "pad": fw.numpy(pad), "tr": fw.numpy(tr), "sq": fw.numpy(sq), "dx": _grad(fw, x),
            "p1": fw.numpy(p1), "dx1": _grad(fw, x1)}


def case_batchnorm2d(fw, seed=8):
    rng = np.random.RandomState(seed)
    np.random.seed(seed)
    bn = fw.place(fw.nn.BatchNorm2d(6))
    x = fw.tensor(rng.normal(2, 3, (5, 6, 4, 4)).astype(np.float32), True)
    bn.train()
    y = bn(x)
    _weighted_sum(fw, y, rng).backward()
    out = {"y": fw.numpy(y), "dx": _grad(fw, x), "dgamma": _grad(fw, bn.weight), "dbeta": _grad(fw, bn.bias),
           "running_mean": fw.numpy(bn.running_mean), "running_var": fw.numpy(bn.running_var)}
    bn.eval()
    out["y_eval"] = fw.numpy(bn(fw.tensor(fw.numpy(x))))
    return out


def case_batchnorm1d(fw, seed=9):
    rng = np.random.RandomState(seed)
    bn = fw.place(fw.nn.BatchNorm1d(7))
    x = fw.tensor(rng.normal(-1, 2, (12, 7)).astype(np.float32), True)
    y = bn(x)
    _weighted_sum(fw, y, rng).backward()
    return {"y": fw.numpy(y), "dx": _grad(fw, x), "dgamma": _grad(fw, bn.weight), "dbeta": _grad(fw, bn.bias),
            "running_mean": fw.numpy(bn.running_mean), "running_var": fw.numpy(bn.running_var)}


# ----------------------------------------------------------------------------- model cases
def _mnist_cnn(fw):
    nn = fw.nn

    class CNN(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(1, 8, 3)
            self.conv2 = nn.Conv2d(8, 16, 3)
            self.l1 = nn.Linear(400, 10)

        def forward(self, x):
            x = x.reshape(shape=[-1, 1, 28, 28])
            x = self.conv1(x).relu().max_pool2d()
            x = self.conv2(x).relu().max_pool2d()
            x = x.flatten()
            return self.l1(x).log_softmax()
    return CNN()


def _make_optimizer(fw, name, params):
    if name == "sgd":
        return fw.optim.SGD(params, lr=0.01, momentum=0)
    if name == "sgdm":
        return fw.optim.SGD(params, lr=0.01, momentum=0.9)
    if name == "adam":
        return fw.optim.Adam(params, lr=1e-3)
    if name == "adamw":
        return fw.optim.AdamW(params, lr=1e-3, weight_decay=1e-2)
    raise ValueError(name)


def _train(fw, model, opt_name, X, Y, steps, criterion):
    fw.place(model)
    opt = _make_optimizer(fw, opt_name, model.parameters())
    rec = {}
    history = []
    for step in range(steps):
        xb, yb = fw.tensor(X), fw.tensor(Y)
        out = model(xb)
        opt.zero_grad()
        loss = criterion(out, yb)
        loss.backward()
        if step == 0:
            rec["out"] = fw.numpy(out)
            rec["loss"] = fw.numpy(loss)
            for i, p in enumerate(model.parameters()):
                rec[f"grad{i}"] = fw.numpy(p.grad)
        opt.step()
        history.append(fw.numpy(loss))
    rec["loss_last"] = fw.numpy(loss)
    if steps > 2:
        rec["loss_history"] = np.concatenate([np.asarray(h, dtype=np.float64).reshape(-1) for h in history])
    for i, p in enumerate(model.parameters()):
        rec[f"param{i}"] = fw.numpy(p)
    return rec


def case_mnist_cnn(fw, opt_name, batch=16, steps=2, seed=0):
    """BASELINE config 0 (README MNIST CNN, NLLLoss) at a fixture-sized batch."""
    np.random.seed(seed)
    model = _mnist_cnn(fw)
    X = np.random.rand(batch, 784).astype(np.float32)
    Y = np.random.randint(0, 10, batch).astype(np.float32)
    return _train(fw, model, opt_name, X, Y, steps, fw.nn.NLLLoss())


def case_vgg_bn(fw, opt_name="adam", batch=8, steps=2, seed=0, widths=(8, 16), hw=8):
    """BASELINE config 2 (VGG-style conv-BN-ReLU stages + FC, log_softmax + NLLLoss) shrunk."""
    nn = fw.nn
    np.random.seed(seed)
    layers, c = [], 3
    for w in widths:
        layers += [nn.Conv2d(c, w, 3, 1, 1), nn.BatchNorm2d(w), nn.ReLU(),
                   nn.Conv2d(w, w, 3, 1, 1), nn.BatchNorm2d(w), nn.ReLU(), nn.MaxPool2d(2)]
        c = w
    layers += [nn.Flatten(), nn.Linear(widths[-1] * (hw // 4) ** 2, 32), nn.ReLU(), nn.Linear(32, 10)]
    body = nn.Sequential(*layers)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.body = body

        def forward(self, x):
            return self.body(x).log_softmax()

    X = np.random.randn(batch, 3, hw, hw).astype(np.float32)
    Y = np.random.randint(0, 10, batch).astype(np.float32)
    return _train(fw, Net(), opt_name, X, Y, steps, nn.NLLLoss())


def case_attention(fw, seq=16, d=8, steps=2, seed=0):
    """BASELINE config 3: single-head self-attention composed from matmul / log_softmax / exp,
    MSELoss, AdamW (examples/attention.py in the reference is an empty stub)."""
    nn = fw.nn
    np.random.seed(seed)

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            mk = lambda: fw.Tensor((np.random.randn(d, d) / np.sqrt(d)).astype(np.float32), requires_grad=True,
                                   is_parameter=True)
            self.wq, self.wk, self.wv, self.wo = mk(), mk(), mk(), mk()

        def forward(self, x):
            q, k, v = x @ self.wq, x @ self.wk, x @ self.wv
            s = (q @ k.T) * (1.0 / np.sqrt(d))
            p = s.log_softmax().exp()
            return (p @ v) @ self.wo

    X = np.random.randn(seq, d).astype(np.float32)
    T = np.random.randn(seq, d).astype(np.float32)
    return _train(fw, Attn(), "adamw", X, T, steps, nn.MSELoss())


def case_tinynet_step(fw, opt_name, seed=0):
    """The reference's optimizer test net (tests/test_step.py:40-51): one step."""
    nn = fw.nn
    rng = np.random.RandomState(seed)
    x0 = rng.randn(1, 3).astype(np.float32)
    w0 = rng.randn(3, 3).astype(np.float32)
    m0 = rng.randn(1, 3).astype(np.float32)

    class TinyNet(nn.Module):
        def __init__(self):
            super().__init__()
            self.x = fw.Tensor(x0.copy(), requires_grad=True)
            self.W = fw.Tensor(w0.copy(), requires_grad=True)
            self.m = fw.Tensor(m0.copy())

        def forward(self):
            out = (self.x @ self.W).relu()
            out = out.log_softmax()
            return out.__mul__(self.m).__add__(self.m).sum()

    net = fw.place(TinyNet().train())
    kinds = {"sgd": ("SGD", {"lr": 0.001, "momentum": 0}), "sgdm": ("SGD", {"lr": 0.001, "momentum": 0.9}),
             "adam": ("Adam", {}), "adamw1": ("AdamW", {"lr": 1e-3, "weight_decay": 1e-1}),
             "adamw2": ("AdamW", {"lr": 1e-3, "weight_decay": 1e-2}), "adamw3": ("AdamW", {"lr": 1e-3, "weight_decay": 1e-3})}
    cls, kwargs = kinds[opt_name]
    opt = getattr(fw.optim, cls)([net.x, net.W], **kwargs)
    out = net.forward()
    out.backward()
    opt.step()
    return {"x": fw.numpy(net.x), "W": fw.numpy(net.W), "out": fw.numpy(out)}


def case_linear_stack(fw, seed=0):
    """tests/test_module.py:32-34: Linear(256,128)-LeakyReLU-Linear(128,1)-ReLU on (8,256)."""
    nn = fw.nn
    np.random.seed(seed)
    model = fw.place(nn.Sequential(nn.Linear(256, 128), nn.LeakyReLU(), nn.Linear(128, 1), nn.ReLU()))
    x = fw.tensor(np.random.randn(8, 256).astype(np.float32), True)
    out = model(x)
    out.sum().backward()
    rec = {"out": fw.numpy(out), "dx": _grad(fw, x)}
    for i, p in enumerate(model.parameters()):
        rec[f"grad{i}"] = fw.numpy(p.grad)
    return rec


def case_activation_mlp(fw, opt_name="sgdm", batch=24, steps=2, seed=0):
    """SURVEY §8-f row 2 layers in one net: LeakyReLU (module.py:744-772), Swish with its learnt beta
    (:818-841), Mish (:844-864), Dropout (mask sampled once, :627-659), MSELoss (:689-712)."""
    nn = fw.nn
    np.random.seed(seed)
    model = nn.Sequential(nn.Linear(20, 32), nn.LeakyReLU(0.1), nn.Linear(32, 32), nn.Swish(), nn.Dropout(0.3),
                          nn.Linear(32, 16), nn.Mish())
    X = np.random.randn(batch, 20).astype(np.float32) * 2
    T = np.random.randn(batch, 16).astype(np.float32)
    model.train()
    return _train(fw, model, opt_name, X, T, steps, nn.MSELoss())


def all_cases():
    """name -> (function, kwargs).  Names double as golden file names."""
    cases = {}
    for k in range(2, 6):
        for s in range(1, 5):
            cases[f"conv2d_k{k}_s{s}"] = (case_conv2d, dict(k=k, s=s, loc=30.0 if (k + s) % 2 else 0.0))
            cases[f"conv1d_k{k}_s{s}"] = (case_conv1d, dict(k=k, s=s, loc=30.0 if (k + s) % 2 else 0.0))
    cases["matmul"] = (case_matmul, {})
    for p in range(1, 5):
        cases[f"maxpool2d_p{p}"] = (case_pool2d, dict(kind="max", p=p))
        cases[f"avgpool2d_p{p}"] = (case_pool2d, dict(kind="avg", p=p))
    for p in range(1, 4):
        cases[f"maxpool1d_p{p}"] = (case_pool1d, dict(kind="max", p=p))
        cases[f"avgpool1d_p{p}"] = (case_pool1d, dict(kind="avg", p=p))
    cases["logsoftmax_nll"] = (case_logsoftmax_nll, {})
    cases["elementwise"] = (case_elementwise, {})
    cases["reductions"] = (case_reductions, {})
    cases["shapes"] = (case_shapes, {})
    cases["batchnorm2d"] = (case_batchnorm2d, {})
    cases["batchnorm1d"] = (case_batchnorm1d, {})
    for o in ("sgd", "sgdm", "adam", "adamw"):
        cases[f"mnist_cnn_{o}"] = (case_mnist_cnn, dict(opt_name=o))
    # multi-step trajectories: bound how far rounding differences are amplified over 10 optimizer steps
    cases["mnist_cnn_adam_10steps"] = (case_mnist_cnn, dict(opt_name="adam", steps=10))
    cases["mnist_cnn_sgdm_10steps"] = (case_mnist_cnn, dict(opt_name="sgdm", steps=10))
    cases["vgg_wide_adam_8steps"] = (case_vgg_bn, dict(opt_name="adam", widths=(32, 64), hw=16, batch=16, steps=8))
    cases["vgg_bn_adam"] = (case_vgg_bn, dict(opt_name="adam"))
    cases["vgg_bn_sgd"] = (case_vgg_bn, dict(opt_name="sgd"))
    # shapes that the tcgen05 TF32 path accepts (channel counts multiple of 32)
    # (16x16 images, batch 16: BatchNorm statistics over >= 256 values, so the comparison measures the
    # kernels and not the conditioning of a 32-sample variance estimate)
    cases["vgg_wide_adam"] = (case_vgg_bn, dict(opt_name="adam", widths=(32, 64), hw=16, batch=16))
    cases["vgg_wide_sgd"] = (case_vgg_bn, dict(opt_name="sgd", widths=(32, 64), hw=16, batch=16))
    cases["conv2d_wide_c32_k64"] = (case_conv2d_wide, dict(c=32, k=64, hw=16, n=4, pad=0))
    cases["conv2d_wide_c64_k32_p1"] = (case_conv2d_wide, dict(c=64, k=32, hw=8, n=3, pad=1))
    cases["conv2d_wide_c128_k128_p1"] = (case_conv2d_wide, dict(c=128, k=128, hw=12, n=2, pad=1))
    # channel counts that are multiples of 64: the layers the FP16x3 tensor-core engine takes (the first conv,
    # 3 -> 64, stays on the direct / SIMT kernels).  With 130-260 k BatchNorm outputs per layer some land within
    # 1e-6 of the ReLU kink or of another element of their max-pool window, where a rounding difference flips
    # relu'(z) / the window's argmax and moves whole gradient entries by 1e-3 on ANY float32 implementation
    # (measured: seed 0 at a ReLU kink with the 3xTF32 engine, seed 6 at a 9e-7 pool near-tie with both
    # tensor-core engines, while the kernels replayed on the recorded operands are within 4e-6 of float64:
    # tools/conv_replay.py).  Seed 23 is the seed in 0..23 that keeps the tensor-core-fed BatchNorm outputs
    # furthest from both discontinuities (>= 7.5e-6, CPU oracle).
    cases["vgg_w64_sgd"] = (case_vgg_bn, dict(opt_name="sgd", widths=(64, 128), hw=16, batch=16, seed=23))
    cases["vgg_w64_adam_8steps"] = (case_vgg_bn, dict(opt_name="adam", widths=(64, 128), hw=16, batch=16, steps=8, seed=23))
    cases["conv2d_wide_c64_k128_p1"] = (case_conv2d_wide, dict(c=64, k=128, hw=12, n=3, pad=1))
    cases["conv2d_wide_c192_k64"] = (case_conv2d_wide, dict(c=192, k=64, hw=9, n=2, pad=0))
    cases["attention_adamw"] = (case_attention, {})
    for o in ("sgd", "sgdm", "adam", "adamw1", "adamw2", "adamw3"):
        cases[f"tinynet_{o}"] = (case_tinynet_step, dict(opt_name=o))
    cases["linear_stack"] = (case_linear_stack, {})
    cases["activation_mlp_sgdm"] = (case_activation_mlp, dict(opt_name="sgdm"))
    cases["activation_mlp_adam"] = (case_activation_mlp, dict(opt_name="adam"))
    return cases


# integer-valued / selection results that must match bit-exactly: one-hot labels, max-pool outputs (a
# maximum of float32 inputs is one of them, whatever the evaluation order)
EXACT_KEYS = {"logsoftmax_nll": ["onehot"]}
# gradients whose *support* must match bit-exactly: max-pool's dx is mask * g / count with a dense random
# g, so (dx != 0) is Max.backward's equality mask (ops_cpu.py:121-127) — the north-star's "argmax indices"
MASK_KEYS = {}
for _p in range(1, 5):
    EXACT_KEYS[f"maxpool2d_p{_p}"] = ["y"]
    MASK_KEYS[f"maxpool2d_p{_p}"] = ["dx"]
for _p in range(1, 4):
    EXACT_KEYS[f"maxpool1d_p{_p}"] = ["y"]
    MASK_KEYS[f"maxpool1d_p{_p}"] = ["dx"]

# Multi-step cases: gradients/outputs of step 0 keep the one-step bar; the end of the trajectory
# (loss_history, loss_last, post-step parameters) is compared at this looser bar because every
# implementation — the float64-drifting oracle included — amplifies 1e-7 rounding differences
# through 8-10 optimizer steps (Adam divides by sqrt(v) ~ |g|).
# Measured on the FP32 SIMT kernels: 7e-6 / 4e-7 / 1.0e-4 (profiles/README.md keeps the GPU numbers).
# vgg_w64_adam_8steps (130-260 k BatchNorm outputs per layer): 2e-4 .. 4e-4 in the three engines, and 1.9e-3 on the
# SIMT kernels after nothing but a different split of BatchNorm's reduction (another summation order moves one
# element across a ReLU / max-pool discontinuity within the eight steps), hence the 3e-3 bar.
TRAJECTORY_RTOL = {"mnist_cnn_adam_10steps": 2e-4, "mnist_cnn_sgdm_10steps": 1e-4, "vgg_wide_adam_8steps": 1e-3,
                   "vgg_w64_adam_8steps": 3e-3}
