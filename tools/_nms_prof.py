import sys, torch
sys.path.insert(0, '.')
import sfod_b200
from sfod_b200 import ops, synth
b, s = synth.boxes_high_suppression(6000, 5)
b = b.cuda(); s = s.cuda()
for _ in range(6):
    k = ops.nms(b, s, 0.7)
torch.cuda.synchronize(); print(len(k))
