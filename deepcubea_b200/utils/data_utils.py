"""stdout tee used by the astar CLI (utils/data_utils.py:12-23)."""
import sys


class Logger(object):
    def __init__(self, filename: str, mode: str = "a"):
        self.terminal = sys.stdout
        self.log = open(filename, mode)

    def write(self, message):
        self.terminal.write(message)
        self.log.write(message)
        self.log.flush()

    def flush(self):
        self.terminal.flush()
