"""Stub of `gym` exposing just the spaces the reference's actor/critic wrappers type-check against.
Test infrastructure only."""
from . import spaces  # noqa: F401
from .spaces import Box, Discrete, Dict, Space  # noqa: F401
