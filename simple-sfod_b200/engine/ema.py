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


def _owners(model: nn.Module, sd: Dict[str, Tensor]):
    """For every state_dict entry: the owning ``_parameters`` / ``_buffers`` dict, the attribute name, the tensor object and
    its address.  ``None`` if an entry is not a plain parameter/buffer (then the caller keeps the slow path)."""
    out = []
    for key, value in sd.items():
        prefix, _, name = key.rpartition(".")
        try:
            mod = model.get_submodule(prefix) if prefix else model
        except AttributeError:
            return None
        for d in (mod._parameters, mod._buffers):
            x = d.get(name)
            if x is not None and x.data_ptr() == value.data_ptr():
                out.append((d, name, x, x.data_ptr()))
                break
        else:
            return None
    return out


class TeacherEMA:
    """Caches the chunk plan for a (student, teacher) model pair; ``step(keep_rate)`` is one kernel launch.

    The key matching of reference :585-601 (two ``state_dict()`` calls, ~0.6 ms of host time for the VGG detector) runs when
    the plan is built; later steps only validate it: every cached tensor must still BE the attribute of its owning module
    (``reset_bn_stats`` replaces the running statistics by new Parameters, reference base.py:318-323) and still live at the
    cached address (``module.to()`` / ``param.data = ...`` move storages).  Any mismatch -- or a module with state that is not
    a plain parameter/buffer -- rebuilds the plan through the reference's matching, including its missing-key exception."""

    def __init__(self, model: nn.Module, model_teacher: nn.Module, world_size: int = 1):
        self.model, self.model_teacher, self.world_size = model, model_teacher, world_size
        self._plan = None
        self._sig = None
        self._watch = None   # [(owner dict, name, tensor, data_ptr)] over student + teacher entries

    def _signature(self, pairs):
        return tuple((k, s.data_ptr(), t.data_ptr(), t.numel()) for k, s, t in pairs)

    def invalidate(self) -> None:
        self._watch = None

    def _still_valid(self) -> bool:
        w = self._watch
        return w is not None and all(d.get(n) is x and x.data_ptr() == p for d, n, x, p in w)

    def _rebuild(self) -> None:
        s_sd, t_sd = self.model.state_dict(), self.model_teacher.state_dict()
        pairs = match_state_dicts(s_sd, t_sd, self.world_size > 1)
        sig = self._signature(pairs)
        if sig != self._sig:
            self._plan = ops.EmaPlan([(s, t) for _, s, t in pairs])
            self._sig = sig
        ws, wt = _owners(self.model, s_sd), _owners(self.model_teacher, t_sd)
        self._watch = None if ws is None or wt is None else ws + wt

    def step(self, keep_rate: float = 0.9996) -> None:
        if not self._still_valid():
            self._rebuild()
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
