"""CPU tests of the student-side training glue: Matcher / subsample_labels /
add_ground_truth_to_proposals / RPN labelling + losses / proposal labelling + Fast R-CNN losses of the plugin classes
against (i) torchvision's own Matcher (independent implementation of the same rule) and (ii) the oracle restatement.
These paths are plain torch (no kernel of the library), so they run without a GPU; the random permutations are injected."""
import math

import pytest
import torch
from torchvision.models.detection._utils import Matcher as TvMatcher

from oracle import d2_cpu as o
import sfod_b200  # noqa: F401
from sfod_b200 import config, modeling
from sfod_b200.modeling import matcher as M
from sfod_b200.structures import Boxes, Instances, pairwise_iou
from sfod_b200.utils.events import EventStorage


def _boxes(n, g, w=1200, h=600):
    xy = torch.rand(n, 2, generator=g) * torch.tensor([w * 0.8, h * 0.8])
    wh = torch.rand(n, 2, generator=g) * torch.tensor([w * 0.3, h * 0.4]) + 8
    return torch.cat([xy, xy + wh], 1)


def _identity_perm(n, device=None):
    return torch.arange(n, device=device)


@pytest.fixture(autouse=True)
def _deterministic_sampling(monkeypatch):
    """subsample_labels draws torch.randperm; inject the identity permutation on both sides."""
    orig = M.subsample_labels
    monkeypatch.setattr(M, "subsample_labels", lambda *a, **k: orig(*a, **{**k, "randperm": _identity_perm}))
    import sfod_b200.modeling.proposal_generator as pg
    import sfod_b200.modeling.roi_heads as rh
    monkeypatch.setattr(pg, "subsample_labels", M.subsample_labels)
    monkeypatch.setattr(rh, "subsample_labels", M.subsample_labels)
    yield


@pytest.mark.parametrize("thr,labels,low", [([0.3, 0.7], [0, -1, 1], True), ([0.5], [0, 1], False)])
def test_matcher_matches_torchvision_and_oracle(thr, labels, low):
    g = torch.Generator().manual_seed(1)
    gt, pr = _boxes(7, g), _boxes(400, g)
    pr[:7] = gt + torch.randn(7, 4, generator=g)          # some high-IoU predictions
    mqm = pairwise_iou(Boxes(gt), Boxes(pr))
    assert torch.equal(mqm, o.pairwise_iou(gt, pr))
    idx, lab = M.Matcher(thr, labels, low)(mqm)
    oi, ol = o.matcher(mqm, thr, labels, low)
    assert torch.equal(idx, oi) and torch.equal(lab, ol)
    tv = TvMatcher(thr[-1], thr[0], allow_low_quality_matches=low)(mqm)      # -1: below low, -2: between, >= 0: matched
    want = torch.where(tv >= 0, torch.tensor(1), torch.where(tv == -2, torch.tensor(-1), torch.tensor(0))).to(torch.int8)
    if len(thr) == 1:
        want = torch.where(tv >= 0, torch.tensor(1), torch.tensor(0)).to(torch.int8)
    assert torch.equal(lab, want)
    assert torch.equal(idx[tv >= 0], tv[tv >= 0])
    # no ground truth: everything is background (label[0]), index 0
    idx0, lab0 = M.Matcher(thr, labels, low)(torch.zeros(0, 5))
    assert idx0.tolist() == [0] * 5 and lab0.tolist() == [labels[0]] * 5


def test_subsample_labels_and_gt_append():
    lab = torch.tensor([1, 0, 0, -1, 1, 0, 1, 1])
    pos, neg = M.subsample_labels(lab, 4, 0.5, 0)
    assert pos.tolist() == [0, 4] and neg.tolist() == [1, 2]
    pos, neg = M.subsample_labels(lab, 100, 0.25, 0)                       # fewer candidates than requested
    assert pos.tolist() == [0, 4, 6, 7] and neg.tolist() == [1, 2, 5]
    p = Instances((100, 100)); p.proposal_boxes = Boxes(torch.zeros(3, 4)); p.objectness_logits = torch.zeros(3)
    out = M.add_ground_truth_to_proposals([Boxes(torch.ones(2, 4))], [p])[0]
    assert len(out) == 5 and out.objectness_logits[3].item() == pytest.approx(math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10))))
    with pytest.raises(ValueError):
        M.add_ground_truth_to_proposals([], [p])


def _cpu_model():
    cfg = config.vgg_source_free_cfg(); cfg.MODEL.DEVICE = "cpu"
    torch.manual_seed(0)
    return modeling.SourceFreeAdaptiveTeacherGeneralizedRCNN(cfg)


def test_rpn_labels_and_losses_match_oracle():
    rpn = _cpu_model().proposal_generator
    g = torch.Generator().manual_seed(2)
    anchors = o.grid_anchors((18, 37), 32, o.generate_cell_anchors())
    gts = [_boxes(5, g), torch.zeros(0, 4), _boxes(2, g)]
    insts = []
    for b in gts:
        i = Instances((600, 1200)); i.gt_boxes = Boxes(b); i.gt_classes = torch.zeros(len(b), dtype=torch.int64); insts.append(i)
    labels, mboxes = rpn.label_and_sample_anchors([Boxes(anchors)], insts)
    ol, ob = o.rpn_label_and_sample_anchors(anchors, gts, 256, 0.5, randperm=lambda n: torch.arange(n))
    for a, b, c, d in zip(labels, ol, mboxes, ob):
        assert torch.equal(a, b) and torch.equal(c, d)
    assert all((l == 1).sum() <= 128 and (l >= 0).sum() <= 256 for l in labels) and (labels[1] == 1).sum() == 0
    logits = torch.randn(3, anchors.shape[0], generator=g); deltas = torch.randn(3, anchors.shape[0], 4, generator=g) * 0.1
    with EventStorage() as st:
        got = rpn.losses([Boxes(anchors)], [logits], labels, [deltas], mboxes)
    want = o.rpn_losses(anchors, logits, ol, deltas, ob, 256)
    for k in want:
        assert torch.allclose(got[k], want[k], rtol=1e-6, atol=1e-7), k
    assert st.latest()["rpn/num_pos_anchors"] == sum(int((l == 1).sum()) for l in labels) / 3


def test_proposal_labelling_and_fast_rcnn_losses_match_oracle():
    model = _cpu_model()
    heads = model.roi_heads
    g = torch.Generator().manual_seed(3)
    props, targets, oprops, otargets = [], [], [], []
    for n_gt in (4, 0):
        pb = _boxes(600, g); gt = _boxes(n_gt, g); gc = torch.randint(0, 8, (n_gt,), generator=g)
        if n_gt:
            pb[:20] = gt.repeat(5, 1) + torch.randn(20, 4, generator=g) * 2
        lg = torch.randn(600, generator=g)
        p = Instances((600, 1200)); p.proposal_boxes = Boxes(pb); p.objectness_logits = lg
        t = Instances((600, 1200)); t.gt_boxes = Boxes(gt); t.gt_classes = gc
        props.append(p); targets.append(t)
        oprops.append(dict(image_size=(600, 1200), proposal_boxes=pb, objectness_logits=lg)); otargets.append(dict(gt_boxes=gt, gt_classes=gc))
    with EventStorage() as st:
        sampled = heads.label_and_sample_proposals(props, targets, branch="supervised_target")
    want = o.label_and_sample_proposals(oprops, otargets, 8, 512, 0.25, True, randperm=lambda n: torch.arange(n))
    for s, w in zip(sampled, want):
        assert torch.equal(s.proposal_boxes.tensor, w["proposal_boxes"]) and torch.equal(s.gt_classes, w["gt_classes"])
        assert torch.equal(s.gt_boxes.tensor, w["gt_boxes"]) and len(s) <= 512
    assert "roi_head/num_target_fg_samples_supervised_target" in st.latest()
    R = sum(len(s) for s in sampled)
    scores = torch.randn(R, 9, generator=g); deltas = torch.randn(R, 32, generator=g) * 0.1
    with EventStorage():
        got = heads.box_predictor.losses((scores, deltas), sampled)
    ref = o.fast_rcnn_losses(scores, deltas, want, 8)
    for k in ref:
        assert torch.allclose(got[k], ref[k], rtol=1e-6, atol=1e-7), k
    empty = heads.box_predictor.losses((torch.zeros(0, 9), torch.zeros(0, 32)), [])
    assert empty["loss_cls"].item() == 0 and empty["loss_box_reg"].item() == 0
