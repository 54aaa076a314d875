import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from nanograd_b200 import backend as B
import nanograd_b200.nn.module as nn
import nanograd_b200.optim.optimizer as optim
from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor
from nanograd_b200.data import BatchLoader
B.init()
np.random.seed(0)
model = bench.build_model(nn); model.gpu()
opt = optim.Adam(model.parameters(), lr=1e-3); crit = nn.NLLLoss()
rng = np.random.RandomState(1)
Xd = rng.randn(2048, 3, 32, 32).astype(np.float32); Yd = rng.randint(0, 10, 2048).astype(np.float32)
x_pin, y_pin = B.pinned_empty((512, 3, 32, 32)), B.pinned_empty((512,))
x_pin[...] = Xd[:512]; y_pin[...] = Yd[:512]
for _ in range(5):
    l = bench.train_step(model, opt, crit, Tensor(x_pin, device=Device.GPU), Tensor(y_pin, device=Device.GPU)); float(l.numpy()[0])
import gc; gc.collect(); gc.disable()
def run(kind, K=20):
    tn = te = tw = 0.0
    t00 = time.perf_counter()
    if kind == "loader":
        it = iter(BatchLoader(Xd, Yd, 512, K, device=Device.GPU))
    for i in range(K):
        t0 = time.perf_counter()
        if kind == "loader":
            xb, yb = next(it)
        else:
            xb, yb = Tensor(x_pin, device=Device.GPU), Tensor(y_pin, device=Device.GPU)
        t1 = time.perf_counter()
        l = bench.train_step(model, opt, crit, xb, yb)
        t2 = time.perf_counter()
        float(l.numpy()[0])
        t3 = time.perf_counter()
        tn += t1 - t0; te += t2 - t1; tw += t3 - t2
    tot = time.perf_counter() - t00
    print(kind, "ms/step %.3f  next %.3f enqueue %.3f wait %.3f" % (tot / K * 1e3, tn / K * 1e3, te / K * 1e3, tw / K * 1e3))
for kind in ("direct", "loader", "direct", "loader"):
    run(kind)
