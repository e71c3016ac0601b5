"""Import shim: the package directory is named ``lora-sdr_b200`` (task layout),
which is not a valid Python identifier, so ``import lora_sdr_b200`` lands here
and this module replaces itself with the real package."""
import importlib.util as _u
import pathlib as _p
import sys as _s

_init = _p.Path(__file__).resolve().parent / "lora-sdr_b200" / "__init__.py"
_spec = _u.spec_from_file_location("lora_sdr_b200", _init, submodule_search_locations=[str(_init.parent)])
_mod = _u.module_from_spec(_spec)
_s.modules["lora_sdr_b200"] = _mod
_spec.loader.exec_module(_mod)
