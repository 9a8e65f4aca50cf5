"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference (adriente/espm v1.1.3).

Only usable where ``/root/reference`` exists (the build container, never the GPU box).
It is used by ``oracle/gen_golden.py`` to (re)generate ``tests/golden/*.npz`` and by the CPU
tests that pin ``oracle/smooth_nmf_oracle.py`` against the real reference when it is present.

The reference's hot path (espm/estimators/{base,smooth_nmf,updates,dicotomy,surrogates}.py,
espm/measures.py, espm/utils.py) depends only on NumPy/SciPy/scikit-learn, but ``espm/conf.py:2``
imports ``exspy._misc.eds.ffast_mac`` and ``espm/utils.py:8`` imports ``exspy.material``; exspy is
not installed here.  We register empty stand-ins for those two imports (no arithmetic lives in
them) so that the reference modules import and run unchanged.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ESPM_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "espm", "estimators"))


def _stub(name, **attrs):
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave like a package
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def load_reference():
    """Import the reference's hot-path modules; returns a namespace with them."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    try:
        import exspy  # noqa: F401  (real package present: nothing to stub)
    except Exception:
        _stub("exspy")
        _stub("exspy._misc")
        _stub("exspy._misc.eds")
        _stub("exspy._misc.eds.ffast_mac", ffast_mac={})

        def _missing(*a, **k):
            raise RuntimeError("exspy is not installed (stubbed for the hot-path oracle)")

        _stub("exspy.material", atomic_to_weight=_missing, density_of_mixture=_missing)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import espm.conf as conf
    import espm.estimators as estimators
    import espm.estimators.updates as updates
    import espm.estimators.dicotomy as dicotomy
    import espm.estimators.surrogates as surrogates
    import espm.measures as measures
    import espm.utils as utils
    import espm.models.base as models_base

    ns = types.SimpleNamespace(
        conf=conf,
        estimators=estimators,
        updates=updates,
        dicotomy=dicotomy,
        surrogates=surrogates,
        measures=measures,
        utils=utils,
        models_base=models_base,
        SmoothNMF=estimators.SmoothNMF,
    )
    return ns
