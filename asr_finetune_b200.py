"""Import shim: the package directory is `asr-finetune_b200/` (not a valid module name); this makes
`import asr_finetune_b200` load it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "asr-finetune_b200")
_spec = importlib.util.spec_from_file_location(
    "asr_finetune_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["asr_finetune_b200"] = _mod
_spec.loader.exec_module(_mod)
