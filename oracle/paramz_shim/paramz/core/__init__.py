"""TEST-ONLY paramz.core stand-in (import-time names only)."""
from . import index_operations, lists_and_dicts, observable, observable_array, parameter_core, pickleable  # noqa: F401
