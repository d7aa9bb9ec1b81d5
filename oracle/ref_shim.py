"""Import shim for the upstream reference (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

The reference (um-dsrg/Super-Resolution-Meta-Attention-Networks) is pure Python and lives at
``/root/reference/Code`` in the build container; it does NOT exist on the GPU box.  This module is
used by ``oracle/make_golden.py`` (to generate the committed fixtures in ``tests/golden/``) and by
the ``-m "not gpu"`` tests that cross-check the restatement against the live reference *when it is
present* (they skip otherwise).

Three import problems on Python 3.12 are patched, all outside the hot path
(``Code/sr_tools/helper_functions.py:5,12,16``): ``collections.Callable``, ``colorama``, ``moviepy``.
"""
import collections
import collections.abc
import os
import sys
import types

REFERENCE_CODE = os.environ.get("DFIR_REFERENCE_CODE", "/root/reference/Code")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_CODE, "SISR", "models"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


_ARCH = None


def import_reference_architectures():
    """Returns the reference module ``SISR.models.attention_manipulators.architectures`` under a
    private name so that it can coexist with this repo's own API-identical ``SISR`` package."""
    global _ARCH
    if _ARCH is not None:
        return _ARCH
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_CODE)
    sys.dont_write_bytecode = True
    collections.Callable = collections.abc.Callable
    _stub("colorama", init=lambda *a, **k: None,
          Fore=types.SimpleNamespace(RED="", GREEN="", RESET="", YELLOW="", BLUE=""))
    for n in ("moviepy", "moviepy.video", "moviepy.video.io", "moviepy.video.io.ImageSequenceClip"):
        _stub(n)
    # the reference and the new repo both ship a top-level `SISR` package: import the reference one
    # in isolation, then stash its modules under a `_ref.` prefix and restore sys.modules.
    saved = {k: v for k, v in sys.modules.items() if k == "SISR" or k.startswith("SISR.")
             or k == "sr_tools" or k.startswith("sr_tools.")}
    for k in saved:
        del sys.modules[k]
    saved_path = list(sys.path)
    sys.path.insert(0, REFERENCE_CODE)
    try:
        import importlib
        arch = importlib.import_module("SISR.models.attention_manipulators.architectures")
        handlers = importlib.import_module("SISR.models.attention_manipulators.handlers")
        advanced = importlib.import_module("SISR.models.advanced.architectures")
        ref_mods = {k: v for k, v in sys.modules.items() if k == "SISR" or k.startswith("SISR.")
                    or k == "sr_tools" or k.startswith("sr_tools.")}
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k == "SISR" or k.startswith("SISR.") or k == "sr_tools" or k.startswith("sr_tools."):
                del sys.modules[k]
        sys.modules.update(saved)
    for k, v in ref_mods.items():
        sys.modules["_ref." + k] = v
    arch._ref_handlers = handlers
    arch._ref_advanced = advanced
    _ARCH = arch
    return arch
