"""CPU-oracle workers of the exp-flip-rate measurement (test infrastructure; also used by bench.py's cpu_baseline leg).

The CUDA path defines ``exp`` / ``softmax`` operation by operation (correctly rounded fp32 exp, sequential fp32 sum) while the
reference's CPU path runs ATen's vectorised ``exp`` / ``softmax`` (MKL VML / Sleef: <= 1 ulp, ISA- and position-dependent, not
reproducible on another machine bit for bit).  A 1-ulp difference in a box coordinate or a score can flip an IoU / threshold
decision.  These workers evaluate the ATen-faithful oracle (and, for a subset, the defined-arithmetic oracle) for a list of
seeds, in a process pool over the host cores; tests/test_gpu_flip_rate.py compares them with the CUDA results.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import torch

RPN_CASES = {  # name -> (synth cfg name, delta std)
    "V_low": ("V", 0.5), "V_high": ("V", 0.05), "R_low": ("R101", 0.5), "R_high": ("R101", 0.05),
}
FRCNN_CASES = ("random", "clustered")
IMAGE = (600, 1200)


def frcnn_inputs(kind: str, seed: int, R: int = 2000, K: int = 8):
    """Fast R-CNN head outputs + proposals.  'random': logits N(0,1.5), deltas N(0,1), uniformly random proposals;
    'clustered': 200 object centres x ~10 jittered proposals each, every cluster confident in one class -- the high-suppression
    distribution a trained detector produces."""
    from sfod_b200 import synth
    if kind == "random":
        cls, dl = synth.box_head_outputs(R, K, seed, 1.5, 1.0)   # top-100 scores straddle the 0.8 pseudo-label threshold
        props = synth.random_rois(1, R, seed + 1)[:, 1:].contiguous()
        return cls, dl, props
    g = torch.Generator().manual_seed(seed)
    clusters = 200
    ctr = torch.rand(clusters, 2, generator=g) * torch.tensor([IMAGE[1], IMAGE[0]], dtype=torch.float32)
    wh = torch.rand(clusters, 2, generator=g) * 200 + 20
    base = torch.cat([ctr - wh / 2, ctr + wh / 2], dim=1)
    which = torch.randint(0, clusters, (R,), generator=g)
    props = base[which] + torch.randn(R, 4, generator=g) * 6.0
    props = torch.stack([torch.minimum(props[:, 0], props[:, 2]), torch.minimum(props[:, 1], props[:, 3]),
                         torch.maximum(props[:, 0], props[:, 2]) + 1.0, torch.maximum(props[:, 1], props[:, 3]) + 1.0], dim=1).contiguous()
    cls_of = torch.randint(0, K, (clusters,), generator=g)
    cls = torch.randn(R, K + 1, generator=g)
    cls[torch.arange(R), cls_of[which]] += torch.rand(R, generator=g) * 3.2
    dl = torch.randn(R, 4 * K, generator=g) * 0.3
    return cls.contiguous(), dl.contiguous(), props


def _rpn_one(case: str, seed: int, exact: bool) -> Dict:
    from oracle import d2_cpu as o
    from sfod_b200 import synth
    cfg_name, std = RPN_CASES[case]
    cfg = getattr(synth, cfg_name)
    logits, deltas, cell, anchors = synth.rpn_head_outputs(cfg, 1, seed, std)
    out = {}
    r = o.rpn_predict_proposals([anchors], [logits], [deltas], [IMAGE], 0.7, 12000, 2000, 0.0, False)[0]
    out["aten_src"] = r["src_index"]
    out["aten_boxes"] = r["proposal_boxes"]
    if exact:
        r = o.rpn_predict_proposals([anchors], [logits], [deltas], [IMAGE], 0.7, 12000, 2000, 0.0, False, exp=o.exp_correctly_rounded)[0]
        out["exact_src"] = r["src_index"]
    return out


def _frcnn_one(kind: str, seed: int, exact: bool) -> Dict:
    from oracle import d2_cpu as o
    cls, dl, props = frcnn_inputs(kind, seed)
    out = {}
    r = o.box_predictor_inference(cls, dl, [props], [IMAGE], 0.05, 0.5, 100)[0]
    out["aten"] = {k: r[k] for k in ("kept_rows", "pred_classes", "scores", "pred_boxes")}
    if exact:
        r = o.box_predictor_inference(cls, dl, [props], [IMAGE], 0.05, 0.5, 100, exp=o.exp_correctly_rounded, softmax=o.softmax_defined)[0]
        out["exact"] = {k: r[k] for k in ("kept_rows", "pred_classes", "scores", "pred_boxes")}
    return out


def work(job: Tuple[str, str, int, bool]) -> Tuple[str, str, int, Dict]:
    kind, case, seed, exact = job
    torch.set_num_threads(1)
    fn = _rpn_one if kind == "rpn" else _frcnn_one
    return kind, case, seed, fn(case, seed, exact)


def run_jobs(jobs: List[Tuple[str, str, int, bool]], workers: int | None = None) -> Dict:
    """Evaluates ``jobs`` in a spawn-context process pool (safe next to an initialised CUDA context)."""
    import multiprocessing as mp
    workers = workers or max(1, min(len(jobs), (os.cpu_count() or 2)))
    results = {}
    if workers == 1:
        for j in jobs:
            k, c, s, r = work(j)
            results[(k, c, s)] = r
        return results
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        for k, c, s, r in pool.imap_unordered(work, jobs, chunksize=2):
            results[(k, c, s)] = r
    return results
