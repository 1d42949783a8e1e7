"""Import shim: the product package lives in the directory
``infrared-colorization-with-resnet-generator-and-patchgan_b200/`` (not a valid Python
identifier); this module loads it under the importable name ``irc_b200``."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "infrared-colorization-with-resnet-generator-and-patchgan_b200")
_spec = importlib.util.spec_from_file_location(
    "irc_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["irc_b200"] = _mod
_spec.loader.exec_module(_mod)
