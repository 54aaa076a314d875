"""Layer-by-layer forward comparison of the vgg_bn test net between two math modes, then all gradients (diagnostic).
python tools/layer_diff.py [modeA modeB]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from nanograd_b200 import backend  # noqa: E402
from nanograd_b200.device import Device  # noqa: E402


def build(fw, widths, hw):
    nn = fw.nn
    np.random.seed(0)
    layers, c = [], 3
    for w in widths:
        layers += [nn.Conv2d(c, w, 3, 1, 1), nn.BatchNorm2d(w), nn.ReLU(),
                   nn.Conv2d(w, w, 3, 1, 1), nn.BatchNorm2d(w), nn.ReLU(), nn.MaxPool2d(2)]
        c = w
    layers += [nn.Flatten(), nn.Linear(widths[-1] * (hw // 4) ** 2, 32), nn.ReLU(), nn.Linear(32, 10)]
    return layers


def run(mode, widths=(64, 128), hw=16, batch=16, fused=False):
    backend.set_math_mode(mode)
    fw = H.device_framework(Device.GPU)
    layers = build(fw, widths, hw)
    X = np.random.randn(batch, 3, hw, hw).astype(np.float32)
    Y = np.random.randint(0, 10, batch).astype(np.float32)
    body = fw.nn.Sequential(*layers)
    fw.place(body)
    x = fw.tensor(X)
    acts = []
    if fused:
        x = body(x)
        acts.append(fw.numpy(x))
    else:
        for l in layers:
            x = l(x)
            acts.append(fw.numpy(x))
    loss = fw.nn.NLLLoss()(x.log_softmax(), fw.tensor(Y))
    loss.backward()
    grads = [fw.numpy(p.grad) for p in body.parameters()]
    return acts, grads, [type(l).__name__ for l in layers]


def rel(a, b):
    s = float(np.abs(b).max())
    return float(np.abs(a.astype(np.float64) - b).max()) / max(s, 1e-300), s


def main():
    H.use_cuda_library()
    a, b = (sys.argv[1:3] + ["simt", "tf32x3"])[:2] if len(sys.argv) >= 3 else ("simt", "tf32x3")
    for fused in (False, True):
        A, GA, names = run(a, fused=fused)
        Bv, GB, _ = run(b, fused=fused)
        print("fused Sequential" if fused else "layer by layer", a, "vs", b)
        for i, (x, y) in enumerate(zip(A, Bv)):
            e, s = rel(y, x.astype(np.float64))
            print(f"  act {i:2d} {names[i] if not fused else 'body':12s} err {e:.1e} (max {s:.1e})")
        for i, (x, y) in enumerate(zip(GA, GB)):
            e, s = rel(y, x.astype(np.float64))
            if s > 1e-12:
                print(f"  grad {i:2d} err {e:.1e} (max {s:.1e}) shape {x.shape}")


if __name__ == "__main__":
    main()
