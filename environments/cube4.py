"""Alias of deepcubea_b200.environments.cube4 (the reference has this environment only in C++, cpp/environments.cpp:262-370)."""
from deepcubea_b200.environments.cube4 import *  # noqa: F401,F403
from deepcubea_b200.environments.cube4 import Cube4, Cube4State  # noqa: F401
