"""Kernel micro-benchmarks (BASELINE.json configs[1]): ROIAlignV2 + NMS sweep, EMA, sort, BN statistics.
CUDA-event timing on the launching stream, L2 flushed between timed iterations, >= 3 warm-ups.
Prints one JSON line per case with algorithmic GB/s against MEASURED_PEAKS.json."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torchvision

import sfod_b200  # noqa: E402
from sfod_b200 import ops, synth  # noqa: E402


def peak_gbs() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def time_fn(fn, iters=10, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def report(name, us_med, us_min, alg_bytes=None, **kw):
    rec = dict(case=name, us_median=round(us_med, 2), us_min=round(us_min, 2))
    if alg_bytes:
        rec["alg_MB"] = round(alg_bytes / 1e6, 2)
        rec["GBps"] = round(alg_bytes / us_med / 1e3, 1)
        rec["frac_of_measured_peak"] = round(alg_bytes / us_med / 1e3 / peak_gbs(), 3)
    rec.update(kw)
    print(json.dumps(rec), flush=True)


def bench_roi(cfg, N, R_per, tv=True):
    dev = "cuda"
    x = synth.features(cfg, N, 1).to(dev)
    xcl = x.contiguous(memory_format=torch.channels_last)
    logits, deltas, cell, anchors = synth.rpn_head_outputs(cfg, N, 2)
    boxes, lg, src, cnt, _ = ops.rpn_select(logits.to(dev), deltas.to(dev), [cfg["image"]] * N, cell_anchors=cell,
                                            feat_hw=(cfg["H"], cfg["W"]), stride=cfg["stride"], post_nms_topk=R_per)
    rois = ops.convert_boxes_to_roi_format([boxes[i] for i in range(N)])
    R = rois.shape[0]
    C, H, W = cfg["C"], cfg["H"], cfg["W"]
    alg = 4 * (N * C * H * W + 5 * R + 49 * R * C)
    sc = 1.0 / cfg["stride"]
    m, mn = time_fn(lambda: ops.roi_align(xcl, rois, (7, 7), sc, 0, True))
    report(f"roi_align_fwd sep NHWC {cfg['name']} N={N} R={R}", m, mn, alg)
    m, mn = time_fn(lambda: ops.roi_align(x, rois, (7, 7), sc, 0, True))
    report(f"roi_align_fwd sep NCHW-in (+transpose) {cfg['name']} N={N} R={R}", m, mn, alg)
    m, mn = time_fn(lambda: ops.roi_align(x, rois, (7, 7), sc, 0, True, exact=True))
    report(f"roi_align_fwd exact {cfg['name']} N={N} R={R}", m, mn, alg)
    if tv:
        m, mn = time_fn(lambda: torchvision.ops.roi_align(x, rois, (7, 7), sc, 0, True))
        report(f"torchvision roi_align_fwd {cfg['name']} N={N} R={R}", m, mn, alg)
    # backward with the student's 512 rois / image
    Rb = 512 * N
    rb = rois[torch.randperm(R, device=dev)[:Rb]].contiguous()
    g = torch.randn(Rb, C, 7, 7, device=dev)
    algb = 4 * (49 * Rb * C + 5 * Rb + N * C * H * W)
    xg = xcl.clone().requires_grad_(True)
    y = ops.roi_align(xg, rb, (7, 7), sc, 0, True)
    m, mn = time_fn(lambda: torch.autograd.grad(y, xg, g, retain_graph=True))
    report(f"roi_align_bwd sep NHWC {cfg['name']} N={N} R={Rb}", m, mn, algb,
           note="through torch.autograd.grad with a synchronize per iteration: includes the host's launch latency (host-bound below ~100 us)")
    # the same C-ABI call enqueued back to back (flush, event, call, event; no synchronize in between): device time only
    from sfod_b200 import _lib
    L = _lib.lib()
    gin = torch.empty((N, C, H, W), device=dev).contiguous(memory_format=torch.channels_last)
    ws = torch.empty(L.sfod_roi_align_bwd_workspace_bytes(N, C, H, W, 1), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def call():
        _lib.check(L.sfod_roi_align_bwd(g.data_ptr(), rb.data_ptr(), N, C, H, W, Rb, 7, 7, sc, 0, 1, gin.data_ptr(), 1, ws.data_ptr(), ws.numel(), st))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    report(f"roi_align_bwd C-ABI queued NHWC {cfg['name']} N={N} R={Rb}", ts[len(ts) // 2], ts[0], algb,
           note="device time: flush, event, C-ABI call (memset + kernel), event enqueued back to back, no synchronize in between")
    if tv:
        xg2 = x.clone().requires_grad_(True)
        y2 = torchvision.ops.roi_align(xg2, rb, (7, 7), sc, 0, True)
        m, mn = time_fn(lambda: torch.autograd.grad(y2, xg2, g, retain_graph=True))
        report(f"torchvision roi_align_bwd {cfg['name']} N={N} R={Rb}", m, mn, algb)


def bench_nms():
    dev = "cuda"
    for kind in ("low", "high"):
        for n in (2000, 4000, 6000, 9990, 12000):
            if kind == "low":
                cfg = synth.V if n <= 9990 else synth.R101
                b, s = synth.boxes_low_suppression(cfg, n, 1234)
            else:
                b, s = synth.boxes_high_suppression(n, 1234)
            bd, sd = b.to(dev), s.to(dev)
            nn = b.shape[0]
            k = ops.nms(bd, sd, 0.7).numel()
            m, mn = time_fn(lambda: ops.nms(bd, sd, 0.7))
            report(f"nms {kind} n={nn} kept={k}", m, mn, 20 * nn + 8 * k, mask_MB=round(nn * ((nn + 63) // 64) * 8 / 1e6, 1))
            m, mn = time_fn(lambda: torchvision.ops.nms(bd, sd, 0.7))
            report(f"torchvision nms {kind} n={nn}", m, mn, 20 * nn + 8 * k)


def bench_rpn_frcnn():
    dev = "cuda"
    for cfg, N in ((synth.V, 1), (synth.V, 8), (synth.R101, 1), (synth.R101, 8)):
        logits, deltas, cell, anchors = synth.rpn_head_outputs(cfg, N, 3)
        ld, dd = logits.to(dev), deltas.to(dev)
        hwa = logits.shape[1]
        f = lambda: ops.rpn_select(ld, dd, [cfg["image"]] * N, cell_anchors=cell, feat_hw=(cfg["H"], cfg["W"]), stride=cfg["stride"])
        m, mn = time_fn(f)
        report(f"rpn_select {cfg['name']} N={N} HWA={hwa}", m, mn, N * (20 * hwa + 20 * 2000), us_per_image=round(m / N, 1))
    for N in (1, 8):
        for std in (4.0, 0.05):
            R = 2000 * N
            cls, dl = synth.box_head_outputs(R, 8, 4, std, 0.5)
            props = synth.random_rois(1, R, 5)[:, 1:].contiguous()
            c, d, p = cls.to(dev), dl.to(dev), props.to(dev)
            f = lambda: ops.frcnn_postprocess(c, d, p, [2000] * N, [(600, 1200)] * N)
            m, mn = time_fn(f)
            report(f"frcnn_postprocess N={N} logit_std={std}", m, mn, 4 * R * (4 + 32 + 9) + N * 100 * 28, us_per_image=round(m / N, 1))


def vgg_state_shapes():
    shapes = []
    cin = 3
    for v in [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512, 'M']:
        if v == 'M':
            continue
        shapes += [(v, cin, 3, 3), (v,), (v,), (v,), (v,), (v,), ()]  # conv w,b ; bn w,b,rm,rv,nbt
        cin = v
    shapes += [(512, 512, 3, 3), (512,), (15, 512, 1, 1), (15,), (60, 512, 1, 1), (60,)]  # rpn head
    shapes += [(1024, 25088), (1024,), (1024, 1024), (1024,), (9, 1024), (9,), (32, 1024), (32,)]  # box head
    shapes += [(512, 512, 3, 3), (512,), (128, 512, 3, 3), (128,), (128, 128, 3, 3), (128,), (1, 128, 3, 3), (1,)]  # D_img
    shapes += [(1024, 2048), (1024,), (1024, 1024), (1024,), (1, 1024), (1,)]  # D_ins (approx.)
    return shapes


def bench_ema():
    dev = "cuda"
    shapes = vgg_state_shapes()
    st, te = [], []
    for s in shapes:
        if s == ():
            st.append(torch.tensor(100, dtype=torch.int64, device=dev)); te.append(torch.tensor(100, dtype=torch.int64, device=dev))
        else:
            st.append(torch.randn(s, device=dev)); te.append(torch.randn(s, device=dev))
    plan = ops.EmaPlan(list(zip(st, te)))
    n = plan.numel
    m, mn = time_fn(lambda: plan.step(0.9996))
    report(f"ema_multi_tensor VGG-like state ({len(shapes)} tensors, {n} elems)", m, mn, 12 * n)
    fl_s = [t for t in st if t.dtype == torch.float32]; fl_t = [t for t in te if t.dtype == torch.float32]

    def foreach():
        torch._foreach_mul_(fl_t, 0.9996)
        torch._foreach_add_(fl_t, fl_s, alpha=1 - 0.9996)
    m, mn = time_fn(foreach)
    report("torch._foreach_mul_/_foreach_add_ EMA (library bar)", m, mn, 12 * n)

    def ref_loop():
        for s_, t_ in zip(st, te):
            t_.copy_(s_ * (1 - 0.9996) + t_ * 0.9996)
    m, mn = time_fn(ref_loop)
    report("reference-style per-tensor EMA loop on GPU", m, mn, 12 * n)


def bench_bn():
    dev = "cuda"
    for (N, C, H, W) in ((8, 64, 600, 1200), (8, 128, 300, 600), (8, 256, 150, 300), (8, 512, 75, 150), (8, 512, 37, 75)):
        for fmt in (torch.contiguous_format, torch.channels_last):
            x = torch.randn(N, C, H, W, device=dev).contiguous(memory_format=fmt)
            w = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
            rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev); nbt = torch.zeros((), dtype=torch.int64, device=dev)
            y = torch.empty_like(x)
            name = "nhwc" if fmt == torch.channels_last else "nchw"
            m, mn = time_fn(lambda: ops.bn_train_forward(x, w, b, rm, rv, nbt, compute_output=False), iters=5)
            report(f"bn_stats {name} {N}x{C}x{H}x{W}", m, mn, 4 * x.numel())
            m, mn = time_fn(lambda: ops.bn_train_forward(x, w, b, rm, rv, nbt, fuse_relu=True, inplace=True), iters=5)
            report(f"bn_stats+apply(relu,inplace) {name} {N}x{C}x{H}x{W}", m, mn, 12 * x.numel())
            m, mn = time_fn(lambda: torch.nn.functional.batch_norm(x, rm, rv, w, b, True, 0.1, 1e-5), iters=5)
            report(f"torch/cuDNN batch_norm train fwd {name} {N}x{C}x{H}x{W}", m, mn, 12 * x.numel())
            del x, y


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    sel = set(a.only.split(",")) if a.only else None
    print(json.dumps(dict(gpu=torch.cuda.get_device_name(0), peak_gbs=peak_gbs())))
    if not sel or "roi" in sel:
        bench_roi(synth.V, 1, 2000); bench_roi(synth.V, 8, 2000); bench_roi(synth.R101, 1, 2000, tv=True); bench_roi(synth.R101, 8, 2000, tv=False)
    if not sel or "nms" in sel:
        bench_nms()
    if not sel or "det" in sel:
        bench_rpn_frcnn()
    if not sel or "ema" in sel:
        bench_ema()
    if not sel or "bn" in sel:
        bench_bn()
