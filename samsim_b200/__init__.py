"""samsim_b200 -- B200-native batch engine for the SAMSIM column timestep.

The product is the CUDA library behind include/samsim_b200.h (samsim_b200/csrc).  This package is
the thin Python host layer over that C ABI: `api.Engine` (ctypes) and `grotz` (the batched mirror of
the reference's grotz(testcase, description) entry point).  There is no CPU fallback: every compute
entry point fails loudly when the CUDA library or a GPU is missing.
"""
from .api import Engine, Config, SamsimError, lib_path, load_library  # noqa: F401

__all__ = ["Engine", "Config", "SamsimError", "lib_path", "load_library"]
