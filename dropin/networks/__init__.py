"""Import shim for the reference's top-level package `networks` (see dropin/_shim.py): modules defined here come from
wsl4mis_b200, everything else from the reference's own code/networks/ directory."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
try:
    from _shim import extend, install_val_2D
finally:
    sys.path.pop(0)
extend(__path__, "networks")
install_val_2D()
