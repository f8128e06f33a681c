"""Environment registry: name -> Environment (the reference's utils/env_utils.py:6-28 for the environments of the hot path)."""
from __future__ import annotations

import math
import re
from typing import Callable, List, Tuple

from ..environments.environment_abstract import Environment


def _cube3(_m) -> Environment:
    from ..environments.cube3 import Cube3
    return Cube3()


def _cube4(_m) -> Environment:
    from ..environments.cube4 import Cube4                   # C++-only in the reference (parallel_weighted_astar.cpp:386)
    return Cube4()


def _n_puzzle(m) -> Environment:
    from ..environments.n_puzzle import NPuzzle
    tiles = int(m.group(1))                                  # "puzzle15" -> 15 tiles -> 4 x 4 board
    return NPuzzle(int(math.sqrt(tiles + 1)))


def _lights_out(m) -> Environment:
    from ..environments.lights_out import LightsOut
    return LightsOut(int(m.group(1)))


_REGISTRY: List[Tuple[str, Callable]] = [(r"^cube3$", _cube3), (r"^cube4$", _cube4), (r"puzzle(\d+)", _n_puzzle), (r"lightsout(\d+)", _lights_out)]


def get_environment(env_name: str) -> Environment:
    key = env_name.lower()
    for pattern, factory in _REGISTRY:
        hit = re.search(pattern, key)
        if hit:
            return factory(hit)
    if key == "sokoban":
        raise ValueError("sokoban is outside the B200 hot path (cube3, cube4, puzzle15/24/35/48, lightsout7)")
    raise ValueError("No known environment %s" % env_name)
