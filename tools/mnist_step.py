"""A few eager README-CNN training steps (BASELINE config 0) for `ncu` launch lists:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/mnist_step.py [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import nanograd_b200.nn.module as nn  # noqa: E402
import nanograd_b200.optim.optimizer as optim  # noqa: E402
from nanograd_b200 import backend as B  # noqa: E402
from nanograd_b200.device import Device  # noqa: E402
from nanograd_b200.tensor import Tensor  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
np.random.seed(0)
model = helpers.build_mnist_cnn(nn).gpu()
opt = optim.SGD(model.parameters(), lr=0.01, momentum=0)
crit = nn.NLLLoss()
xb = Tensor(np.random.rand(128, 784).astype(np.float32), device=Device.GPU)
yb = Tensor(np.random.randint(0, 10, 128).astype(np.float32), device=Device.GPU)
for _ in range(steps):
    out = model(xb)
    opt.zero_grad()
    loss = crit(out, yb)
    loss.backward()
    opt.step()
B.sync()
print("loss", float(loss.numpy()[0]), "launches/step", B.launch_count() / steps)
