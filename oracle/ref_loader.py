"""The UNMODIFIED reference (BaekMS/Sep-TFAnet-VAD) as an importable checker / CPU baseline
(TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__ and bench.py's reference legs import this).

The reference is Python, so there is nothing to compile with gcc. Its "build" is byte-compilation:
``build_ref()`` compiles the handful of reference modules on the inference path - where they lie
under ``/root/reference``, nothing is copied into the repository's history - to sourceless
bytecode files (``*.refbc``: the ``.pyc`` format under another extension, because repository snapshots
commonly drop ``*.pyc``) under ``oracle/_ref/`` (git-ignored, not gpurun-ignored: like a built ``.so``
it travels to the GPU box, where ``/root/reference`` does not exist); a small meta-path finder imports
them under their original module names. ``load()`` imports the reference from
``/root/reference`` when it is present (this container) and from ``oracle/_ref`` otherwise (GPU box).

Two imports of the reference are dead for the arithmetic and absent from this image - ``turtle``
(``model/combined_loss.py:1``) and ``matplotlib`` (``Our_utils/utlis_inference.py:3``); they are
satisfied with empty stub modules.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SEPTFA_REFERENCE", "/root/reference")
REF_BUILT = os.path.join(HERE, "_ref")
EXT = ".refbc"

#: the reference modules on (or next to) the hot path: SURVEY.md section 8(a)
MODULES = (
    "model/model.py",                          # SeparationModel                     A1-A14
    "model/online_class_unknown_targets.py",   # OnlineSaving (unknown targets)      A15
    "model/online_class_known_targets.py",     # OnlineSaving (known targets)        A16
    "model/pit_wrapper.py",                    # PITLossWrapper                      A17
    "model/combined_loss.py",                  # reorder_source_mse, calc_sisdr      A17
    "model/sdr.py",                            # PairwiseNegSDR (test.py's pw_mtx)   (f)3
    "Our_utils/utlis_inference.py",            # save_audio / plot_spectrogram / save_vad   A19
    "Our_utils/utils_test.py",                 # simple VAD from masks (test.py)     (f)3
)


def build_ref(verbose=False):
    """Byte-compile the reference modules into oracle/_ref (no-op when /root/reference is absent)."""
    if not os.path.isdir(REF_SRC):
        return False
    for rel in MODULES:
        src = os.path.join(REF_SRC, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(REF_BUILT, rel[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        if verbose:
            print("compiled", rel, "->", os.path.relpath(dst, HERE))
    with open(os.path.join(REF_BUILT, "BUILT_FROM"), "w") as f:
        f.write(f"{REF_SRC}\npython {sys.version_info.major}.{sys.version_info.minor}\n")
    return True


def available():
    return os.path.isdir(REF_SRC) or os.path.exists(os.path.join(REF_BUILT, "model", "model" + EXT))


class _BuiltFinder(importlib.abc.MetaPathFinder):
    """Imports ``model.*`` / ``Our_utils.*`` from the byte-compiled files under oracle/_ref."""
    PACKAGES = ("model", "Our_utils")

    def find_spec(self, fullname, path=None, target=None):
        parts = fullname.split(".")
        if parts[0] not in self.PACKAGES or len(parts) > 2:
            return None
        if len(parts) == 1:   # the reference's packages have no __init__.py: plain namespace packages
            spec = importlib.machinery.ModuleSpec(fullname, None, is_package=True)
            spec.submodule_search_locations = [os.path.join(REF_BUILT, parts[0])]
            return spec
        fn = os.path.join(REF_BUILT, parts[0], parts[1] + EXT)
        if not os.path.exists(fn):
            return None
        return importlib.util.spec_from_loader(fullname, importlib.machinery.SourcelessFileLoader(fullname, fn))


def _stub_dead_imports():
    if "turtle" not in sys.modules:
        t = types.ModuleType("turtle")
        t.forward = None
        sys.modules["turtle"] = t
    try:
        importlib.import_module("matplotlib.pyplot")
    except Exception:  # noqa: BLE001  (absent in this image)
        mp, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mp.pyplot = pp
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mp, pp


_loaded = None


def load():
    """Import the reference; returns a namespace with ``model`` (model/model.py), ``pit_wrapper``,
    ``combined_loss``, ``OnlineSaving`` (unknown targets), ``online_known`` (module) and ``root``."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = REF_SRC if os.path.isdir(REF_SRC) else REF_BUILT
    if not available():
        raise RuntimeError("the reference is neither at /root/reference nor byte-compiled under oracle/_ref "
                           "(run __graft_entry__.build() in the build container)")
    sys.dont_write_bytecode = True   # /root/reference is read-only
    _stub_dead_imports()
    if root == REF_SRC:
        if root not in sys.path:
            sys.path.insert(0, root)
    elif not any(isinstance(f, _BuiltFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _BuiltFinder())
    ns = types.SimpleNamespace(root=root, kind="source" if root == REF_SRC else "bytecode")
    ns.model = importlib.import_module("model.model")
    ns.pit_wrapper = importlib.import_module("model.pit_wrapper")
    ns.combined_loss = importlib.import_module("model.combined_loss")
    ns.OnlineSaving = importlib.import_module("model.online_class_unknown_targets").OnlineSaving
    ns.online_known = importlib.import_module("model.online_class_known_targets")
    _loaded = ns
    return ns


if __name__ == "__main__":
    print("built" if build_ref(verbose=True) else "reference sources not present; nothing built")
