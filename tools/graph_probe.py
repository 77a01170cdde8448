import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import sfod_b200
from sfod_b200 import config, modeling, engine
from sfod_b200.engine.graph import GraphedTeacherStep
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(42)
cfg = config.vgg_source_free_cfg(); cfg.MODEL.DEVICE = "cuda"
teacher = modeling.SourceFreeAdaptiveTeacherGeneralizedRCNN(cfg).cuda().train()
student = modeling.SourceFreeAdaptiveTeacherGeneralizedRCNN(cfg).cuda().train()
ema = engine.TeacherEMA(student, teacher)
x = torch.randint(0, 256, (2, 3, 600, 1200), dtype=torch.uint8, device="cuda")
try:
    gs = GraphedTeacherStep(teacher, x.shape, lambda: ema.step(0.9996))
    d, p = gs.run(x)
    print("ok", [len(i) for i in d], [len(i) for i in p])
    with torch.no_grad():
        _, _, r = teacher(x, branch="unsup_data_weak")
    print("eager", [len(i) for i in r])
except Exception:
    traceback.print_exc()
