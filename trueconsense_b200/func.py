"""Small CLI helpers with the reference's names (TrueConsense/func.py): the argparse help
formatter the command line uses and the ANSI colour table of its error messages."""
from __future__ import annotations

import argparse
import os
import shutil


class MyHelpFormatter(argparse.RawTextHelpFormatter):
    """Raw-text help whose option column scales with the terminal and whose help strings get a
    "(default: ...)" suffix unless they already mention a default (func.py:6-28)."""

    def __init__(self, prog):
        width = shutil.get_terminal_size().columns
        os.environ["COLUMNS"] = str(width)
        super().__init__(prog, max_help_position=min(max(24, width // 2), 80))

    def _get_help_string(self, action):
        text = action.help
        default = action.default
        if default is not None and default != argparse.SUPPRESS and "default" not in text.lower():
            text += " (default: " + str(default) + ")"
        return text


def _sgr(code: int) -> str:
    return f"\033[{code}m"


class color:
    """ANSI escape sequences under the reference's attribute names (func.py:31-40): ``color.RED + text + color.END``."""


for _name, _code in (("PURPLE", 95), ("CYAN", 96), ("DARKCYAN", 36), ("BLUE", 94), ("GREEN", 92), ("YELLOW", 93),
                     ("RED", 91), ("BOLD", 1), ("UNDERLINE", 4), ("END", 0)):
    setattr(color, _name, _sgr(_code))
del _name, _code
