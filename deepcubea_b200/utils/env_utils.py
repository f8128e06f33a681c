"""Name -> environment registry (utils/env_utils.py:6-28), restricted to the environments of the hot path."""
import math
import re

from ..environments.environment_abstract import Environment


def get_environment(env_name: str) -> Environment:
    name = env_name.lower()
    m = re.search(r"puzzle(\d+)", name)
    if name == "cube3":
        from ..environments.cube3 import Cube3
        return Cube3()
    if m is not None:
        from ..environments.n_puzzle import NPuzzle
        return NPuzzle(int(math.sqrt(int(m.group(1)) + 1)))
    if "lightsout" in name:
        from ..environments.lights_out import LightsOut
        return LightsOut(int(re.search(r"lightsout(\d+)", name).group(1)))
    if name == "sokoban":
        raise ValueError("%s is outside the B200 hot path (cube3, puzzle15/24/35/48, lightsout7)" % env_name)
    raise ValueError("No known environment %s" % env_name)
