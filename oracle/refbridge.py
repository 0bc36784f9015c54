"""Import the real reference (nextpyp/cet_pick) as the ground truth, in THIS container only.

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing on a `-m gpu`
test, smoke() or bench.py path may call this; it is used by tests/golden/make_golden.py and by
the CPU-only tests that are skipped when the reference is absent.

The non-arithmetic third-party modules the reference imports but this image lacks
(progress, mrcfile, sknetwork, matplotlib) are replaced by inert stubs.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("CET_PICK_REF", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "cet_pick"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    """Put the reference on sys.path and stub the missing non-arithmetic imports."""
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    class _Bar:
        def __init__(self, *a, **k): pass
        def next(self): pass
        def finish(self): pass
        suffix = ""

    try:
        import progress.bar  # noqa: F401
    except Exception:
        p = _stub("progress")
        p.bar = _stub("progress.bar", Bar=_Bar)
    try:
        import mrcfile  # noqa: F401
    except Exception:
        class _Mrc:
            def __init__(self, path): self.path, self.data = path, None
            def set_data(self, d): self.data = d
            def __enter__(self): return self
            def __exit__(self, *a):
                import numpy as np
                if self.data is not None:
                    np.save(self.path + ".npy", self.data)
        _stub("mrcfile", new=lambda path, overwrite=True: _Mrc(path), open=None)
    try:
        import sknetwork.topology  # noqa: F401
    except Exception:
        s = _stub("sknetwork")
        s.topology = _stub("sknetwork.topology", get_connected_components=None,
                           get_largest_connected_component=None)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        m = _stub("matplotlib")
        m.pyplot = _stub("matplotlib.pyplot")
        m.use = lambda *a, **k: None


def decode_module():
    install()
    import cet_pick.models.decode as d
    return d


def utils_module():
    install()
    import cet_pick.models.utils as u
    return u


def create_model(arch="unet_4", heads=None, head_conv=32, last_k=3):
    install()
    from cet_pick.models.model import create_model as cm
    return cm(arch, heads if heads is not None else {"hm": 1, "proj": 32}, head_conv, last_k=last_k)
