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


class color:
    PURPLE = "\033[95m"
    CYAN = "\033[96m"
    DARKCYAN = "\033[36m"
    BLUE = "\033[94m"
    GREEN = "\033[92m"
    YELLOW = "\033[93m"
    RED = "\033[91m"
    BOLD = "\033[1m"
    UNDERLINE = "\033[4m"
    END = "\033[0m"
