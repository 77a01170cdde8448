import sys, os, json, torch, torchvision
sys.path.insert(0, '.')
import sfod_b200
from sfod_b200 import config, modeling, ops, synth
sys.path.insert(0, '.')
import bench
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(42)
cfg = config.vgg_source_free_cfg(); cfg.MODEL.DEVICE = "cpu"
teacher = modeling.SourceFreeAdaptiveTeacherGeneralizedRCNN(cfg).cuda().train()
cap = {}
orig = ops.rpn_select
def hook(logits, deltas, image_sizes, **kw):
    cap['logits'] = logits.detach().cpu(); cap['deltas'] = deltas.detach().cpu(); cap['kw'] = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in kw.items()}
    cap['sizes'] = image_sizes
    return orig(logits, deltas, image_sizes, **kw)
ops.rpn_select = hook
import sfod_b200.modeling.proposal_generator as pg
if hasattr(pg, 'ops'): pg.ops.rpn_select = hook
imgs = bench.synth_images(8, 1234).cuda()
with torch.no_grad():
    teacher(imgs, branch="unsup_data_weak")
lg, dl = cap['logits'], cap['deltas']
print('logits stats', float(lg.mean()), float(lg.std()), 'deltas std', float(dl.std()))
from oracle import d2_cpu
kw = cap['kw']
cell = kw['cell_anchors']; H, W = kw['feat_hw']; stride = kw['stride']
anchors = synth.grid_anchors(H, W, stride, cell)
for i in range(2):
    props = d2_cpu.decode_proposals([anchors], [dl[i:i+1]])[0][0]
    b = props.clone(); h, w = cap['sizes'][i]
    b[:, 0::2] = b[:, 0::2].clamp(0, w); b[:, 1::2] = b[:, 1::2].clamp(0, h)
    s = lg[i]
    ok = ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)
    b = b[ok]; s = s[ok]
    order = s.argsort(descending=True, stable=True); b = b[order]; s = s[order]
    keep = torchvision.ops.nms(b, s, 0.7)
    r = int(keep[1999]) if len(keep) >= 2000 else -1
    print('image', i, 'n', len(b), 'kept total', len(keep), 'rank of 2000th kept', r, 'tiles', r // 64 + 1)
