"""TEST INFRASTRUCTURE ONLY -- loader of the real reference PyTAT module built by oracle/Makefile
into oracle/_ref/ (git-ignored, travels to the GPU box).  Returns None when it is not built."""
import glob
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load_reference_tat():
    if "m" in _cache:
        return _cache["m"]
    files = glob.glob(os.path.join(_HERE, "_ref", "TAT*.so"))
    if not files:
        _cache["m"] = None
        return None
    saved = sys.modules.get("TAT")
    try:
        # torch must be imported BEFORE the reference extension: loading them in the other order crashes
        # inside torch's own pybind initialisation (both embed pybind11 internals / libstdc++ symbols)
        import torch  # noqa: F401
        spec = importlib.util.spec_from_file_location("TAT", files[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception:  # missing libopenblas on a foreign box, ABI mismatch, ...
        mod = None
    finally:
        if saved is not None:
            sys.modules["TAT"] = saved
        else:
            sys.modules.pop("TAT", None)
    _cache["m"] = mod
    return mod
