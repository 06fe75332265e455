"""Loader for the unmodified Python reference (only present in the build container).

`/root/reference` does not exist on the GPU box: tests that need it are skipped there
and rely on the committed fixtures under tests/golden/ instead.
"""
import importlib.util
import os

REFERENCE_ROOT = os.environ.get("MSGWAM_REFERENCE", "/root/reference")


def load_reference():
    """Return a *fresh* instance of the reference's lib/libprop.py module, or None."""
    path = os.path.join(REFERENCE_ROOT, "lib", "libprop.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("_msgwam_reference_libprop", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
