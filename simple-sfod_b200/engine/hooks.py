"""detectron2's hook protocol and the minimal loop that drives it, plus hot-path hooks on top of this package's engine.

The reference's trainers are detectron2 ``TrainerBase`` subclasses and its hooks subclass ``detectron2.engine.hooks.HookBase``
(reference daod/engine/hooks/val_loss.py:8,90: ``after_step`` reads ``self.trainer.iter`` / ``max_iter`` / ``storage``).  The
protocol below is detectron2's: ``before_train -> (before_step -> run_step -> after_step)* -> after_train``, hooks get a
weak ``trainer`` attribute when registered, ``trainer.storage`` is an ``EventStorage`` that is current during ``train``.
Only the protocol is provided -- data loading, checkpointing and evaluation hooks are the reference's control plane.
"""
from __future__ import annotations

import weakref
from typing import List, Optional

from ..utils.events import EventStorage


class HookBase:
    """detectron2.engine.HookBase: subclasses override any of the four (five) callbacks."""

    trainer: "TrainerBase" = None

    def before_train(self):
        pass

    def after_train(self):
        pass

    def before_step(self):
        pass

    def after_backward(self):
        pass

    def after_step(self):
        pass

    def state_dict(self):
        return {}


class TrainerBase:
    """detectron2.engine.TrainerBase: iteration bookkeeping + hook dispatch; subclasses implement ``run_step``."""

    def __init__(self) -> None:
        self._hooks: List[HookBase] = []
        self.iter: int = 0
        self.start_iter: int = 0
        self.max_iter: int = 0
        self.storage: Optional[EventStorage] = None

    def register_hooks(self, hooks: List[Optional[HookBase]]) -> None:
        hooks = [h for h in hooks if h is not None]
        for h in hooks:
            assert isinstance(h, HookBase)
            h.trainer = weakref.proxy(self)   # no reference cycle trainer <-> hook, as in detectron2
        self._hooks.extend(hooks)

    def train(self, start_iter: int, max_iter: int) -> None:
        self.iter = self.start_iter = start_iter
        self.max_iter = max_iter
        with EventStorage(start_iter) as self.storage:
            try:
                self.before_train()
                for self.iter in range(start_iter, max_iter):
                    self.before_step()
                    self.run_step()
                    self.after_step()
                self.iter += 1   # detectron2: the final value reflects the number of completed iterations
            finally:
                self.after_train()

    def before_train(self):
        for h in self._hooks:
            h.before_train()

    def after_train(self):
        self.storage.iter = self.iter
        for h in self._hooks:
            h.after_train()

    def before_step(self):
        self.storage.iter = self.iter
        for h in self._hooks:
            h.before_step()

    def after_backward(self):
        for h in self._hooks:
            h.after_backward()

    def after_step(self):
        for h in self._hooks:
            h.after_step()

    def run_step(self):
        raise NotImplementedError


class TeacherEMAHook(HookBase):
    """Mean-teacher update as a hook, in one native launch per update (``keep_rate * teacher + (1 - keep_rate) * student``).

    The reference performs the update inline in ``run_step`` at one of two places, selected here with ``when``:

    * ``"before_step"`` (default) -- daod/engine/trainers/adaptive_teacher.py:215-223: at the START of iteration ``it``, before the
      teacher pseudo-labels that iteration's batch: ``it == BURN_UP_STEP`` copies the student (keep_rate 0 through the same
      formula, also when BURN_UP_STEP is 0), later iterations with ``(it - BURN_UP_STEP) % TEACHER_UPDATE_ITER == 0`` apply
      ``EMA_KEEP_RATE``; nothing happens during burn-in (``it < BURN_UP_STEP``).
    * ``"after_step"`` -- daod/engine/trainers/source_free_adaptive_teacher_single.py:581 (and ``_mosaic``): after the optimizer
      step of EVERY iteration, with the fixed keep rate (0.9996 in those trainers); no burn-in copy.
    """

    def __init__(self, student, teacher, keep_rate: float = 0.9996, period: int = 1, burn_up_step: int = 0, world_size: int = 1,
                 when: str = "before_step"):
        from .ema import TeacherEMA
        if when not in ("before_step", "after_step"):
            raise ValueError(f"when must be 'before_step' or 'after_step', got {when!r}")
        self._ema = TeacherEMA(student, teacher, world_size=world_size)
        self._keep_rate, self._period, self._burn, self._when = float(keep_rate), int(period), int(burn_up_step), when

    def before_step(self):
        if self._when != "before_step":
            return
        it = self.trainer.iter
        if it < self._burn:
            return
        if it == self._burn:
            self._ema.step(0.0)
        elif (it - self._burn) % self._period == 0:
            self._ema.step(self._keep_rate)

    def after_step(self):
        if self._when == "after_step":
            self._ema.step(self._keep_rate)
