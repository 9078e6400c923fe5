"""pytest plugin (``-p tests.ref_alias_plugin``) for the acceptance run of the
REFERENCE's own test files against this package (tests/test_reference_suite.py).

It imports the unmodified reference (baseline/_ref, or /root/reference/src in the
build container) with exactly the two modules of INTEGRATION.md section 1
replaced before anything else binds them:

    openseize.core.numerical  ->  openseize_b200.core.numerical
    openseize.core.producer   ->  openseize_b200.core.producer

so the reference's own L4 operators (``filtering/bases.py``, ``resampling/
resampling.py``, ``spectra/estimators.py``, ``tools/pipeline.py`` -- unmodified)
run over the CUDA hot path.  ``OSZ_REF_FAKE=1`` installs the numpy stand-in
kernels instead (CPU suite: host logic only).  TEST INFRASTRUCTURE ONLY.
"""

import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402


def _stub(name, **attrs):
    try:
        importlib.import_module(name)
        return
    except Exception:
        pass
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    if "." in name:
        parent, child = name.rsplit(".", 1)
        setattr(sys.modules[parent], child, mod)


def _install():
    where = refload.location()
    if where is None:
        raise RuntimeError("reference not found (baseline/_ref or /root/reference)")
    refload.stub_plotting()
    # openseize.demos (imported by the reference's tests/conftest.py) wants these
    _stub("wget")
    _stub("tkinter", Tk=object, __path__=[])
    _stub("tkinter.filedialog")
    _stub("tkinter.messagebox")
    if where not in sys.path:
        sys.path.insert(0, where)
    import openseize_b200.core.numerical as gnm
    import openseize_b200.core.producer as gpro

    sys.modules["openseize.core.producer"] = gpro       # before `import openseize`
    sys.modules["openseize.core.numerical"] = gnm
    import openseize
    import openseize.core

    openseize.core.producer = gpro
    openseize.core.numerical = gnm
    openseize.producer = gpro.producer
    if os.environ.get("OSZ_REF_FAKE") == "1":
        from tests import fake_backend

        fake_backend.install_raw()
    return gnm, gpro


_GNM, _GPRO = _install()


def pytest_sessionfinish(session, exitstatus):
    """The run only counts if the reference's operators really were bound to the
    replacement modules (and, on a GPU, the native library is what ran)."""
    import openseize.filtering.bases as bases
    import openseize.resampling.resampling as rs
    import openseize.spectra.estimators as est

    assert bases.nm is _GNM and est.nm is _GNM, "reference operators not bound to the replacement"
    assert rs.polyphase_resample is _GNM.polyphase_resample
    assert bases.producer is _GPRO.producer and bases.__file__.startswith(refload.location())
    if os.environ.get("OSZ_REF_FAKE") != "1":
        from openseize_b200 import _abi

        assert _abi.launch_count() > 0, "no CUDA kernel launched during the reference's tests"
