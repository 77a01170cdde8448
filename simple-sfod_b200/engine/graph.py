"""CUDA-graph capture of the teacher pseudo-labelling step.

The teacher step of this package is free of host synchronisation between its first and last kernel (the RPN hands a padded
device batch to the ROI heads, the fused post-processing leaves device-side counts; tests/test_gpu_plugins.py checks that a
step performs ONE device->host read).  That makes the whole launch sequence -- preprocess, ~60 cuDNN/cuBLAS calls, ~45 launches
of libsfod_b200, the EMA update -- capturable: it is recorded once into a CUDA graph and replayed with a single host call per
step.  With strict-fp32 library math the step is GPU-bound and the gain is small; with PyTorch's default TF32 convolutions the
step is short enough (~11 ms for 8 images) for launch latency to matter.

What the reference does at this point is a Python loop over images with a ``.item()`` / boolean-mask sync per image
(reference daod/engine/trainers/source_free_adaptive_teacher.py:385-390, 256-280).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
from torch import Tensor

from ..structures import Instances


class GraphedTeacherStep:
    """``run(images)`` = ``teacher(images, branch="unsup_data_weak")`` + ``after()`` (e.g. the EMA launch) replayed from a CUDA
    graph, followed by the one host read of the counts.  ``images``: (N, 3, H, W) uint8 / float32 batch of a fixed shape."""

    def __init__(self, teacher: torch.nn.Module, shape, after: Optional[Callable[[], None]] = None, threshold: float = 0.8,
                 dtype: torch.dtype = torch.uint8, warmup: int = 3):
        dev = next(teacher.parameters()).device
        self.teacher, self.threshold = teacher, threshold
        self.static_in = torch.zeros(tuple(shape), dtype=dtype, device=dev)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        predictor = teacher.roi_heads.box_predictor
        was = predictor.defer_host_read
        predictor.defer_host_read = True                    # the one D2H read of the step moves behind the graph
        try:
            with torch.cuda.stream(side), torch.no_grad():  # warm-up on a side stream: lazy initialisations (cuDNN plans, workspaces)
                for _ in range(warmup):
                    teacher(self.static_in, branch="unsup_data_weak")
                    if after is not None:
                        after()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(self.graph), torch.no_grad():
                _, self._rpn, self._roih = teacher(self.static_in, branch="unsup_data_weak")
                if after is not None:
                    after()
        finally:
            predictor.defer_host_read = was
        self._batch = self._roih[0]._sfod_batch             # padded device-side detections + counts (static buffers of the graph)

    def run(self, images: Tensor) -> Tuple[List[Instances], List[Instances]]:
        """Returns (detections, pseudo-labels) of the batch, like ``process_pseudo_label(teacher(...)[2], thr, "roih",
        "thresholding")``; the tensors are views of the graph's static output buffers (valid until the next ``run``)."""
        self.static_in.copy_(images, non_blocking=True)
        self.graph.replay()
        b = self._batch
        b._host = None                                       # new contents: the counts must be read again (ONE D2H read)
        if b.proposal_batch is not None:
            b.proposal_batch._host = None
        dets, _ = b.instances()
        return dets, b.pseudo_labels()
