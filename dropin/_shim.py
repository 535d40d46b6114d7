"""Shared by the shim packages: make `import utils.metrics` / `networks.efficientunet` / ... keep resolving to the reference's own
files for everything this repo does not replace.

Each shim directory (`utils/`, `networks/`, `dataloaders/`) is a REGULAR package (it has an `__init__.py`), so it wins over
the reference's namespace-package directory of the same name even though the script directory `code/` comes first on
`sys.path` (a regular package found later beats namespace portions found earlier).  `extend(__path__, name)` then appends
the reference's `code/<name>/` directory -- located from `sys.path[0]` / the working directory, i.e. where the unchanged
`train_*.py` script lives -- so sub-modules that exist only there are still importable."""
import os
import sys


def extend(pkg_path, name):
    seen = set(os.path.realpath(p) for p in pkg_path)
    for base in [sys.path[0] if sys.path else "", os.getcwd()] + [p for p in sys.path if p]:
        d = os.path.join(base or os.getcwd(), name)
        if os.path.isdir(d) and os.path.realpath(d) not in seen and not os.path.exists(os.path.join(d, "__init__.py")):
            pkg_path.append(d)
            seen.add(os.path.realpath(d))


def install_val_2D():
    """`val_2D` is a plain module NEXT TO the script, so no path order can shadow it (`sys.path[0]` is the script directory).
    The scripts import it after `dataloaders` / `networks` / `utils` (train_weakly_supervised_pCE_GatedCRFLoss_2D.py:23-28):
    the first shim package to load registers the GPU validation module under that name, and `from val_2D import
    test_single_volume` then finds it in sys.modules.  Set WSL4MIS_KEEP_REFERENCE_VAL=1 to keep the reference's host loop."""
    if os.environ.get("WSL4MIS_KEEP_REFERENCE_VAL", "0") == "1" or "val_2D" in sys.modules:
        return
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("val_2D", os.path.join(here, "val_2D.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["val_2D"] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules["val_2D"]
        raise
