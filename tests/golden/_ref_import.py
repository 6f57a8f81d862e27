"""Import the UNMODIFIED reference model files from /root/reference (container-only helper).

Used by tests/golden/make_golden.py and by the container-only oracle-vs-reference tests.  The
GPU box has no /root/reference: everything that must run there reads the committed fixtures in
tests/golden/ instead.
"""
import os
import sys
import types

REF_ROOT = "/root/reference"
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "live2diff", "animatediff", "models"))


def import_reference_models():
    """Return the reference `live2diff.animatediff.models` sub-modules, imported unmodified.

    `live2diff/__init__.py` pulls in the full pipeline (LCMScheduler, ...), so an empty package
    object whose __path__ points at the reference tree is registered instead.
    """
    if not reference_available():
        raise RuntimeError("reference tree not present (this helper only works in the build container)")
    shim = os.path.abspath(SHIM)
    if shim not in sys.path:
        sys.path.insert(0, shim)
    for name, sub in (("live2diff", "live2diff"), ("live2diff.animatediff", "live2diff/animatediff")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REF_ROOT, sub)]
            sys.modules[name] = pkg
    import importlib

    mods = {}
    for m in ("resnet", "positional_encoding", "attention", "stream_motion_module", "motion_module",
              "unet_blocks_streaming", "unet_depth_streaming", "unet_blocks_warmup", "unet_depth_warmup"):
        mods[m] = importlib.import_module(f"live2diff.animatediff.models.{m}")
    return mods
