"""How often does the CUDA path's defined ``exp`` / ``softmax`` arithmetic change a DISCRETE result relative to the reference's
ATen-CPU arithmetic?  (VERDICT r1 "What's weak" 1.)

For >= 200 seeds x {VGG, R101-C4 anchors} x {N(0,0.5) deltas, high-suppression deltas} the RPN keep sets, and for >= 200 seeds x
{random, clustered "trained-like"} head outputs the Fast R-CNN detection and pseudo-label sets, are computed on the GPU and by
the ATen-faithful oracle (process pool over the host cores, tests/flip_workers.py).  The test

* asserts a stated bound on the flip rate (sets that differ at all / elements that differ),
* checks boxes and scores to 1e-5 relative on the common subset UNCONDITIONALLY (also when a flip occurred),
* checks bit-equality against the defined-arithmetic oracle on every 4th seed (200+ additional bit-exact cases),
* writes the measured rates to gpurun_out/flip_rate.json (copied to profiles/ by the builder).

SFOD_FLIP_SEEDS overrides the number of seeds (default 200).
"""
import json
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import flip_workers as fw  # noqa: E402
import sfod_b200  # noqa: F401,E402
from sfod_b200 import ops, synth  # noqa: E402

pytestmark = pytest.mark.gpu
SEEDS = int(os.environ.get("SFOD_FLIP_SEEDS", "200"))
BATCH = 8
# Stated bounds.  Measured (profiles/r2_flip_rate.json): see DESIGN.md section 3; the bounds leave a factor of ~5.
MAX_SETS_DIFFERING = 0.05      # fraction of images whose keep / detection set differs at all
MAX_ELEMENTS_DIFFERING = 5e-4  # fraction of kept proposals / detections that differ
MAX_PSEUDO_SETS_DIFFERING = 0.02


def _rel_close(a: torch.Tensor, b: torch.Tensor, rtol=1e-5, scale: float = 1200.0) -> bool:
    """|a - b| <= 1e-5 |b| + 1e-6 * scale: 1e-5 relative fp32 (BASELINE.json), with an absolute floor of 1e-6 of the coordinate
    range for box corners that are small differences of large terms (x1 = ctr - w/2: one ulp of w is not relative to x1)."""
    a, b = a.double(), b.double()
    return bool(((a - b).abs() <= rtol * b.abs() + 1e-6 * scale).all())


def _record(name, payload):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "flip_rate.json")
    data = {}
    if os.path.exists(path):
        with open(path) as f:
            data = json.load(f)
    data[name] = payload
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def test_rpn_keep_set_flip_rate(cuda_device):
    jobs = [("rpn", case, 5000 + s, s % 4 == 0) for case in fw.RPN_CASES for s in range(SEEDS)]
    oracle = fw.run_jobs(jobs)
    summary = {}
    for case, (cfg_name, std) in fw.RPN_CASES.items():
        cfg = getattr(synth, cfg_name)
        n_sets = n_kept = n_diff = n_exact = 0
        for s0 in range(0, SEEDS, BATCH):
            seeds = [5000 + s for s in range(s0, min(SEEDS, s0 + BATCH))]
            lg, dl = [], []
            for sd in seeds:
                logits, deltas, cell, anchors = synth.rpn_head_outputs(cfg, 1, sd, std)
                lg.append(logits); dl.append(deltas)
            boxes, _, src, cnt, invalid = ops.rpn_select(torch.cat(lg).to(cuda_device), torch.cat(dl).to(cuda_device), [fw.IMAGE] * len(seeds),
                                                         cell_anchors=cell, feat_hw=(cfg["H"], cfg["W"]), stride=cfg["stride"],
                                                         pre_nms_topk=12000, post_nms_topk=2000, nms_thresh=0.7)
            cnt = cnt.cpu().tolist(); src = src.cpu(); boxes = boxes.cpu()
            for i, sd in enumerate(seeds):
                r = oracle[("rpn", case, sd)]
                got = src[i, :cnt[i]]
                if "exact_src" in r:                       # defined arithmetic: bit-exact, always
                    assert torch.equal(got, r["exact_src"]), (case, sd)
                    n_exact += 1
                a, b = set(got.tolist()), set(r["aten_src"].tolist())
                n_kept += len(b); n_diff += len(a ^ b); n_sets += bool(a ^ b)
                # common subset: same proposal -> same box to 1e-5 (unconditional)
                pos_g = {v: j for j, v in enumerate(got.tolist())}
                common = [v for v in r["aten_src"].tolist() if v in pos_g]
                gi = torch.tensor([pos_g[v] for v in common], dtype=torch.long)
                ri = torch.tensor([j for j, v in enumerate(r["aten_src"].tolist()) if v in pos_g], dtype=torch.long)
                assert _rel_close(boxes[i][gi], r["aten_boxes"][ri]), (case, sd)
        summary[case] = dict(images=SEEDS, sets_differing=n_sets, kept=n_kept, elements_differing=n_diff, bit_exact_vs_defined_oracle=n_exact)
        assert n_sets <= max(1, MAX_SETS_DIFFERING * SEEDS), summary
        assert n_diff <= max(2, MAX_ELEMENTS_DIFFERING * n_kept), summary
    _record("rpn_keep_sets_vs_aten_exp", summary)
    print("RPN flip rate:", summary)


def test_frcnn_detection_and_pseudo_label_flip_rate(cuda_device):
    jobs = [("frcnn", kind, 7000 + s, s % 4 == 0) for kind in fw.FRCNN_CASES for s in range(SEEDS)]
    oracle = fw.run_jobs(jobs)
    summary = {}
    for kind in fw.FRCNN_CASES:
        n_sets = n_det = n_diff = n_pl_sets = n_pl = n_exact = 0
        for s0 in range(0, SEEDS, BATCH):
            seeds = [7000 + s for s in range(s0, min(SEEDS, s0 + BATCH))]
            ins = [fw.frcnn_inputs(kind, sd) for sd in seeds]
            rows = [len(x[2]) for x in ins]
            out = ops.frcnn_postprocess(torch.cat([x[0] for x in ins]).to(cuda_device), torch.cat([x[1] for x in ins]).to(cuda_device),
                                        torch.cat([x[2] for x in ins]).to(cuda_device), rows, [fw.IMAGE] * len(seeds), pseudo_thresh=0.8)
            cnt = out["count"].cpu().tolist(); pc = out["pseudo_count"].cpu().tolist()
            g_rows, g_cls, g_sc, g_bx = out["rows"].cpu(), out["classes"].cpu(), out["scores"].cpu(), out["boxes"].cpu()
            for i, sd in enumerate(seeds):
                r = oracle[("frcnn", kind, sd)]
                k = cnt[i]
                if "exact" in r:
                    e = r["exact"]
                    assert k == len(e["scores"]) and torch.equal(g_rows[i, :k], e["kept_rows"]) and torch.equal(g_cls[i, :k], e["pred_classes"])
                    assert torch.equal(g_sc[i, :k], e["scores"]) and torch.equal(g_bx[i, :k], e["pred_boxes"]), (kind, sd)
                    assert pc[i] == int((e["scores"] > 0.8).sum())
                    n_exact += 1
                ref = r["aten"]
                got_keys = list(zip(g_rows[i, :k].tolist(), g_cls[i, :k].tolist()))
                ref_keys = list(zip(ref["kept_rows"].tolist(), ref["pred_classes"].tolist()))
                a, b = set(got_keys), set(ref_keys)
                n_det += len(b); n_diff += len(a ^ b); n_sets += bool(a ^ b)
                pl_g = set(got_keys[:pc[i]]); npl = int((ref["scores"] > 0.8).sum()); pl_r = set(ref_keys[:npl])
                n_pl += npl; n_pl_sets += bool(pl_g ^ pl_r)
                pos = {v: j for j, v in enumerate(got_keys)}
                ri = [j for j, v in enumerate(ref_keys) if v in pos]
                gi = [pos[ref_keys[j]] for j in ri]
                assert _rel_close(g_bx[i][gi], ref["pred_boxes"][ri]) and _rel_close(g_sc[i][gi], ref["scores"][ri], scale=0.0), (kind, sd)
        summary[kind] = dict(images=SEEDS, detection_sets_differing=n_sets, detections=n_det, detections_differing=n_diff,
                             pseudo_labels=n_pl, pseudo_label_sets_differing=n_pl_sets, bit_exact_vs_defined_oracle=n_exact)
        assert n_pl > 0
        assert n_sets <= max(1, MAX_SETS_DIFFERING * SEEDS), summary
        assert n_diff <= max(2, MAX_ELEMENTS_DIFFERING * n_det), summary
        assert n_pl_sets <= max(1, MAX_PSEUDO_SETS_DIFFERING * SEEDS), summary
    _record("frcnn_detections_vs_aten_exp_softmax", summary)
    print("Fast R-CNN flip rate:", summary)
