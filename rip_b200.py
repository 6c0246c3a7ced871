"""Importable alias for the package directory `opencl-development-real-time-image-processing_b200/`
(its name has hyphens, so `import` cannot spell it).  `import rip_b200 as rip`."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "opencl-development-real-time-image-processing_b200")
_NAME = "opencl_development_real_time_image_processing_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"),
                                                   submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
_mod = sys.modules[_NAME]
globals().update({k: v for k, v in vars(_mod).items() if not k.startswith("__")})
PKG_DIR = _PKG_DIR
