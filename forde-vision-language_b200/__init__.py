"""forde_b200 -- B200-native (sm_100a) implementation of FORDE's fast-loop hot path.

Import as ``forde_b200`` (the directory name carries the reference's repository name and is not a valid
Python identifier; ``forde_b200.py`` at the repository root maps the name onto this package).
"""
from . import _lib  # noqa: F401
from ._lib import FordeError  # noqa: F401

__all__ = ["FordeError"]
