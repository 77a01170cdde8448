"""Small host-side utilities mirroring the detectron2 helpers the plugins call (event storage)."""
from .events import EventStorage, get_event_storage  # noqa: F401
