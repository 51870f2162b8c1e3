"""Load the UNMODIFIED reference (Kyaw-Thiha/animal-vision) in place from /root/reference.

Container-only tooling: used by tools/make_golden.py to generate tests/golden/*.npz and to
cross-check the oracle restatement.  Nothing in tests -m gpu, smoke() or bench.py imports this
module -- /root/reference does not exist on the GPU box.

Three shims are needed to import the reference at all:
  1. animals/__init__.py imports animals/cat.py first, which has unresolved merge-conflict
     markers -> register an empty `animals` package whose __path__ points at the reference.
  2. animals/cat.py: keep the `Tina-animals` side of the conflict (the only runnable side),
     exec the resolved text as module `animals.cat`.
  3. ml/classic_rgb_to_hsi imports `colour` (not installed) and its analytic branch is gated
     on torch.cuda.is_available(); stub `colour` and force the analytic branch onto CPU tensors
     by running the reference's own function with `torch.cuda.is_available` -> True and
     device="cuda" tensors redirected to CPU.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REF = os.environ.get("AVB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "animals"))


def _install_package_stub():
    if "animals" in sys.modules and getattr(sys.modules["animals"], "__avb_stub__", False):
        return
    pkg = types.ModuleType("animals")
    pkg.__path__ = [os.path.join(REF, "animals")]
    pkg.__avb_stub__ = True
    sys.modules["animals"] = pkg
    if "colour" not in sys.modules:
        sys.modules["colour"] = types.ModuleType("colour")
    if REF not in sys.path:
        sys.path.insert(0, REF)


def species(module: str, cls: str):
    """Return the reference class `cls` from animals/<module>.py (e.g. species('dog','Dog'))."""
    _install_package_stub()
    if module == "cat":
        return _cat_class()
    if module == "honeybee":
        _patch_classic_hsi()
    return getattr(importlib.import_module(f"animals.{module}"), cls)


def _cat_class():
    name = "animals.cat"
    if name in sys.modules and hasattr(sys.modules[name], "Cat"):
        return sys.modules[name].Cat
    src = open(os.path.join(REF, "animals", "cat.py")).read().split("\n")
    out, mode = [], "both"
    for line in src:
        if line.startswith("<<<<<<<"):
            mode = "head"
        elif line.startswith("=======") and mode == "head":
            mode = "theirs"
        elif line.startswith(">>>>>>>"):
            mode = "both"
        elif mode != "head":
            out.append(line)
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF, "animals", "cat.py")
    sys.modules[name] = mod
    exec(compile("\n".join(out), mod.__file__, "exec"), mod.__dict__)
    return mod.Cat


_patched = False


def _patch_classic_hsi():
    """Make the reference's analytic ("cuda") branch of classic_rgb_to_hsi run on CPU tensors.

    The function body is executed unmodified; only torch.cuda.is_available() and the
    device="cuda" argument of torch.as_tensor are redirected.
    """
    global _patched
    if _patched:
        return
    import torch

    chsi = importlib.import_module("ml.classic_rgb_to_hsi.classic_rgb_to_hsi")
    orig = chsi.classic_rgb_to_hsi
    if torch.cuda.is_available():
        _patched = True
        return

    class _TorchProxy:
        def __getattr__(self, k):
            return getattr(torch, k)

        class cuda:  # noqa: N801
            @staticmethod
            def is_available():
                return True

        @staticmethod
        def as_tensor(x, dtype=None, device=None):
            return torch.as_tensor(x, dtype=dtype, device="cpu")

    def patched(frame, **kw):
        g = orig.__globals__
        saved = g["torch"]
        g["torch"] = _TorchProxy()
        try:
            return orig(frame, **kw)
        finally:
            g["torch"] = saved

    chsi.classic_rgb_to_hsi = patched
    hb = importlib.import_module("animals.honeybee")
    hb.classic_rgb_to_hsi = patched
    _patched = True


def classic_rgb_to_hsi():
    _install_package_stub()
    _patch_classic_hsi()
    return importlib.import_module("ml.classic_rgb_to_hsi.classic_rgb_to_hsi").classic_rgb_to_hsi


def module(name: str):
    """Import a top-level reference module (uv_helpers, uv_mappers, animals.animal_utils ...)."""
    _install_package_stub()
    return importlib.import_module(name)


def mstpp_module():
    path = os.path.join(REF, "ml/MST_plus_plus/predict_code/architecture/MST_Plus_Plus.py")
    spec = importlib.util.spec_from_file_location("avb_ref_mstpp", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
