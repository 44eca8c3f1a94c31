"""Import the UNMODIFIED reference package (`fullwave`) in an environment that lacks matplotlib.

The reference imports matplotlib at module import time (fullwave/medium.py:8, utils/plot_utils.py)
only for plotting helpers; the image has no matplotlib, so inert stand-in modules are registered
before the import.  Nothing of the reference is modified.  Used by the fixture generators under
tools/ (in the build container, from /root/reference) and by bench.py's reference arm (on the GPU
box, from baseline/_ref).
"""

from __future__ import annotations

import importlib
import importlib.machinery
import sys
import types
from pathlib import Path


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(f"{self.__name__}.{name}")
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return self

    def __or__(self, other):  # `Figure | None` annotations are evaluated at def time
        return self

    __ror__ = __or__

    def __getitem__(self, item):
        return self

    def __mro_entries__(self, bases):  # usable as a base class
        return (object,)


def _stub(name: str) -> None:
    if name in sys.modules:
        return
    m = _Anything(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    sys.modules[name] = m


def import_fullwave(root: str | Path | None = None):
    """root: directory that CONTAINS the `fullwave` package (default: /root/reference, else baseline/_ref)."""
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.colors",
                  "matplotlib.patches", "matplotlib.figure", "matplotlib.axes", "matplotlib.cm",
                  "mpl_toolkits", "mpl_toolkits.axes_grid1"):
            _stub(n)
    if root is None:
        here = Path(__file__).resolve().parent.parent
        for cand in (Path("/root/reference"), here / "baseline" / "_ref"):
            if (cand / "fullwave" / "__init__.py").exists():
                root = cand
                break
    if root is None:
        raise ImportError("reference package not found (neither /root/reference nor baseline/_ref)")
    root = str(root)
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("fullwave")
