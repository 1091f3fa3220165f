"""Mirror of `vipe.ext` for the one operator module this repo implements (vipe/ext/__init__.py:39-46)."""

from . import slam_ext  # noqa: F401

__all__ = ["slam_ext"]
