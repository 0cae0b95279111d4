"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box; nothing in
the ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this at run time.
It is used by ``oracle/gen_golden.py`` (fixture generation) and by the
``not gpu`` tests that validate the numpy restatement against the reference.

The reference's ``poduqnn/podnnmodel.py:6-7`` imports tensorflow and
tensorflow_probability at module top and ``poduqnn/mesh.py:6`` imports meshio;
none of them is installed, and none is touched by the POD / snapshot /
projection path, so they are replaced by ``MagicMock`` stubs.
"""
import importlib
import importlib.util
import os
import sys
from unittest import mock

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "poduqnn"))


def _stub(name):
    if name not in sys.modules:
        sys.modules[name] = mock.MagicMock(name=name)


def load():
    """Return the reference's ``poduqnn`` package (pod, acceleration, podnnmodel, mesh)."""
    if not available():
        raise RuntimeError("reference tree not present (only in the build container)")
    for name in ("tensorflow", "tensorflow_probability", "meshio",
                 "matplotlib", "matplotlib.pyplot", "yaml"):
        try:
            importlib.import_module(name)
        except Exception:  # noqa: BLE001 - absent or broken: stub it
            _stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import poduqnn.pod  # noqa: F401
    import poduqnn.acceleration  # noqa: F401
    import poduqnn.podnnmodel  # noqa: F401
    import poduqnn.mesh  # noqa: F401
    return sys.modules["poduqnn"]


def load_field(case):
    """Load ``u(X, t, mu)`` and ``HP`` from experiments/<case>/hyperparams.py."""
    load()
    path = os.path.join(REFERENCE_ROOT, "experiments", case, "hyperparams.py")
    spec = importlib.util.spec_from_file_location(f"_ref_hp_{case}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.u, mod.HP
