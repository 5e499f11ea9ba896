"""Stub of `yapf` (only used by the reference for pretty-printing configs). Test infrastructure only."""
