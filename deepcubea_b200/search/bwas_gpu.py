"""Batch weighted A* for ONE start state: `BWASGpu`, the single-instance face of the device-driven engine (search/engine.py)."""
from .engine import NONE, BWASGpu, BWASResult, CapacityHeuristic, SearchEngine  # noqa: F401
