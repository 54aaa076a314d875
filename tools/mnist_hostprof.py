import cProfile, pstats, os, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
from nanograd_b200 import backend as B
import nanograd_b200.nn.module as nn
import nanograd_b200.optim.optimizer as optim
from nanograd_b200.device import Device
from nanograd_b200.tensor import Tensor
B.init()
np.random.seed(0)
model = helpers.build_mnist_cnn(nn); model.gpu()
opt = optim.SGD(model.parameters(), lr=0.01, momentum=0)
crit = nn.NLLLoss()
X = np.random.rand(128, 784).astype(np.float32); Y = np.random.randint(0, 10, 128).astype(np.float32)
xb, yb = Tensor(X, device=Device.GPU), Tensor(Y, device=Device.GPU)
def step():
    out = model(xb); opt.zero_grad(); loss = crit(out, yb); loss.backward(); opt.step(); return loss
for _ in range(20): step()
B.sync()
l0 = B.launch_count()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
B.sync()
pr.disable()
print("launches per step", (B.launch_count() - l0) / 200)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
