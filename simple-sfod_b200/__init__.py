"""sfod_b200 -- B200-native implementation of the simple-SFOD teacher-student pseudo-labelling hot path.

The package directory is ``simple-sfod_b200/`` (the name the build contract fixes); it is imported as
``sfod_b200`` through the loader module ``sfod_b200.py`` at the repository root.
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401

__all__ = ["ops"]
