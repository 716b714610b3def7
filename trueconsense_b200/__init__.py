"""trueconsense_b200 — B200-native pileup-and-call hot path with TrueConsense's Python API.

Modules mirror the reference package (``indexing``, ``Coverage``, ``Events``, ``Ambig``,
``Sequences``, ``ORFs``, ``Outputs``, ``TrueConsense``); ``gpu`` is the ctypes binding of the C-ABI
(include/trueconsense_b200.h), ``bamio`` / ``synth`` / ``reads`` the host substrate.
"""
from .version import __version__  # noqa: F401
