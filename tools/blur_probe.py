import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import torchvision.transforms.functional as TF
import sfod_b200
from sfod_b200 import ops
g = torch.Generator().manual_seed(5)
x = torch.randint(0, 256, (2, 3, 40, 50), dtype=torch.uint8, generator=g)
for s in (0.1, 1.0):
    got = ops.gaussian_blur(x.cuda(), [s, s]).cpu()
    ks = 2 * max(1, int(np.ceil(3 * s))) + 1
    want = TF.gaussian_blur(x[0], [ks, ks], [s, s])
    print(s, ks, "x", x[0, 0, :2, :6].tolist(), "got", got[0, 0, :2, :6].tolist(), "want", want[0, :2, :6].tolist())
    print(ops.gaussian_kernel1d(ks, s))
