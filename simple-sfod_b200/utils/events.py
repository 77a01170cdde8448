"""detectron2.utils.events subset: the plugins log scalars through ``get_event_storage().put_scalar`` (reference
daod/modeling/roi_heads/source_free_adaptive_teacher_roi_heads.py:207-213).  With detectron2 installed its storage is
used; otherwise a minimal context-managed storage with the same calls (and a silent default one) stands in."""
from __future__ import annotations

from collections import defaultdict, deque
from typing import Dict, List

_CURRENT: List["EventStorage"] = []


class EventStorage:
    HISTORY_WINDOW = 4096   # per-key window: no writer drains the storage, so it must not grow with the run length

    def __init__(self, start_iter: int = 0):
        self.iter = start_iter
        self._latest: Dict[str, float] = {}
        self._history: Dict[str, deque] = defaultdict(lambda: deque(maxlen=self.HISTORY_WINDOW))

    def put_scalar(self, name: str, value, smoothing_hint: bool = True) -> None:
        value = float(value)
        self._latest[name] = value
        self._history[name].append(value)

    def put_scalars(self, *, smoothing_hint: bool = True, **kwargs) -> None:
        for k, v in kwargs.items():
            self.put_scalar(k, v, smoothing_hint)

    def latest(self) -> Dict[str, float]:
        return dict(self._latest)

    def history(self, name: str) -> List[float]:
        return list(self._history[name])

    def step(self) -> None:
        self.iter += 1

    def __enter__(self):
        _CURRENT.append(self)
        return self

    def __exit__(self, *exc):
        assert _CURRENT[-1] is self
        _CURRENT.pop()


_DEFAULT = EventStorage()


def _resolve_backend():
    """Decided once, at import: a genuine detectron2 storage is used only when detectron2 is importable AND is not this
    package's own shim (whose ``get_event_storage`` IS the function below -- deferring to it would recurse)."""
    import sys
    d2 = sys.modules.get("detectron2")
    if d2 is not None and getattr(d2, "__sfod_shim__", False):
        return None
    try:
        from detectron2.utils.events import get_event_storage as d2_get  # pragma: no cover - detectron2 is absent in the build environment
    except ImportError:
        return None
    return None if getattr(sys.modules.get("detectron2"), "__sfod_shim__", False) else d2_get  # pragma: no cover


_D2_GET = _resolve_backend()


def get_event_storage() -> EventStorage:
    if _D2_GET is not None and _D2_GET is not get_event_storage:  # pragma: no cover
        return _D2_GET()
    return _CURRENT[-1] if _CURRENT else _DEFAULT
