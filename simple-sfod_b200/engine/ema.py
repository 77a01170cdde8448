"""Mean-teacher EMA: ``_update_teacher_model`` of the reference trainers
(reference daod/engine/trainers/source_free_adaptive_teacher.py:583-603, same body adaptive_teacher.py:339-358).

The reference builds a new state_dict with three temporaries per tensor and ``load_state_dict``s it (~36 B/element in
>300 launches).  Here the (student, teacher) tensor pairs are matched once by key -- including the DDP ``module.``
prefix strip of :586-589 and the missing-key exception of :600-601 -- and every step is ONE launch of
``sfod_ema_multi_tensor`` at the algorithmic 12 B/element, bit-identical to the reference's arithmetic.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
from torch import Tensor, nn

from .. import ops


def match_state_dicts(student_sd: Dict[str, Tensor], teacher_sd: Dict[str, Tensor], strip_module_prefix: bool) -> List[Tuple[str, Tensor, Tensor]]:
    """Key matching of reference :585-601.  Pure host logic (no device work)."""
    if strip_module_prefix:
        student_sd = {key[7:]: value for key, value in student_sd.items()}
    pairs = []
    for key, value in teacher_sd.items():
        if key in student_sd.keys():
            pairs.append((key, student_sd[key], value))
        else:
            raise Exception("{} is not found in student model".format(key))
    return pairs


class TeacherEMA:
    """Caches the chunk plan for a (student, teacher) model pair; ``step(keep_rate)`` is one kernel launch."""

    def __init__(self, model: nn.Module, model_teacher: nn.Module, world_size: int = 1):
        self.model, self.model_teacher, self.world_size = model, model_teacher, world_size
        self._plan = None
        self._sig = None

    def _signature(self, pairs):
        return tuple((k, s.data_ptr(), t.data_ptr(), t.numel()) for k, s, t in pairs)

    def step(self, keep_rate: float = 0.9996) -> None:
        pairs = match_state_dicts(self.model.state_dict(), self.model_teacher.state_dict(), self.world_size > 1)
        sig = self._signature(pairs)
        if sig != self._sig:  # storages moved (e.g. reset_bn_stats re-created the running stats): rebuild the plan
            self._plan = ops.EmaPlan([(s, t) for _, s, t in pairs])
            self._sig = sig
        self._plan.step(keep_rate)

    @property
    def numel(self) -> int:
        return 0 if self._plan is None else self._plan.numel


@torch.no_grad()
def update_teacher_model(model: nn.Module, model_teacher: nn.Module, keep_rate: float = 0.9996, world_size: int = 1) -> None:
    """Functional form with the reference's argument meaning: ``trainer._update_teacher_model(keep_rate)`` becomes
    ``update_teacher_model(trainer.model, trainer.model_teacher, keep_rate, comm.get_world_size())``."""
    cache = model_teacher.__dict__.setdefault("_sfod_ema", None)
    if cache is None or cache.model is not model or cache.world_size != world_size:
        cache = TeacherEMA(model, model_teacher, world_size)
        model_teacher.__dict__["_sfod_ema"] = cache
    cache.step(keep_rate)
