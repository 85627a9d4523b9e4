"""Importable alias of the package directory `fpga-cnn-object-detection-accelerator_b200/` (a hyphenated
directory name cannot appear in an `import` statement):  `import fpga_cnn_b200 as fc`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("fpga-cnn-object-detection-accelerator_b200")
