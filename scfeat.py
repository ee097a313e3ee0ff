"""Importable alias for the package directory ``tf-keras-speech-commands_b200`` (its name is not a valid
Python identifier): ``import scfeat`` gives the same module object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module('tf-keras-speech-commands_b200')
sys.modules[__name__] = _pkg
