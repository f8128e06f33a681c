"""`python search_methods/astar.py ...` -- the reference's command line, served by deepcubea_b200."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepcubea_b200.search_methods.astar import *  # noqa: E402,F401,F403
from deepcubea_b200.search_methods.astar import AStar, Node, get_path, main  # noqa: E402,F401

if __name__ == "__main__":
    main()
