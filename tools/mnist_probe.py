import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
from nanograd_b200 import backend as B
B.init()
for mode in ("fp32", "simt", "tf32x3", "fp32"):
    B.set_math_mode(mode)
    t0 = time.time()
    r = bench.mnist_cnn_images_per_sec(B)
    print(mode, r["images_per_sec"], r["ms_per_step"], "wall %.2f s" % (time.time() - t0), B.mem_stats())
