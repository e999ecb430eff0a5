"""Append-to-file logger that also echoes to a stream (robosat/log.py:8-27)."""

import os
import sys


class Log:
    def __init__(self, path, out=sys.stdout):
        self.out = out
        self.fp = open(path, "a")

    def log(self, msg):
        self.fp.write(msg + os.linesep)
        self.fp.flush()
        if self.out:
            print(msg, file=self.out)
