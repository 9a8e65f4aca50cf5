"""Boundary contract of the drop-in that needs no GPU: what the reference's own consumers
check or read on a fitted estimator -- ``isinstance(..., NMFEstimator)`` (eds_spim.py:607, 639), the ``L_`` matrix
(base.py:287-291, utils.py:39-76) and the ``get_losses`` record layout (base.py:479-517)."""
import numpy as np
import pytest

from oracle import ref_import
from oracle import smooth_nmf_oracle as orc

needs_ref = pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("nx,ny", [(2, 2), (3, 7), (4, 2), (9, 5)])
def test_laplacian_object_is_the_reference_matrix(nx, ny):
    """test_laplacian.py:17-31 of the reference compares with the explicit Neumann stencil; so do we."""
    from espm_b200.ops import create_laplacian_matrix
    Lm = create_laplacian_matrix(nx, ny)
    assert Lm.shape == (nx * ny, nx * ny)
    dense = Lm.toarray()
    assert dense.dtype == np.float32
    assert np.array_equal(dense, orc.laplacian_dense(nx, ny))
    rng = np.random.default_rng(nx * 10 + ny)
    H = rng.uniform(size=(3, nx * ny))
    assert np.allclose(H @ Lm, orc.laplacian_apply(H, (nx, ny)), atol=1e-6)       # updates.py:96
    assert np.allclose(Lm @ H.T, orc.laplacian_apply(H, (nx, ny)).T, atol=1e-6)   # measures.py:577
    assert np.allclose(Lm.dot(H.T), Lm @ H.T)
    # largest eigenvalue <= sigmaL = 8 (test_laplacian.py:33-38)
    assert np.max(np.linalg.eigvalsh(dense.astype(np.float64))) <= 8.0 + 1e-6


def test_identity_laplacian_object():
    from espm_b200.ops import GridLaplacian
    Lm = GridLaplacian(None, identity=7)
    assert Lm.shape == (7, 7) and Lm.shape_2d is None
    assert np.array_equal(Lm.toarray(), np.eye(7, dtype=np.float32))


@needs_ref
@pytest.mark.parametrize("nx,ny", [(2, 3), (5, 4), (6, 6)])
def test_laplacian_object_vs_live_reference(nx, ny):
    ref = ref_import.load_reference()
    from espm_b200.ops import create_laplacian_matrix
    theirs = ref.utils.create_laplacian_matrix(nx, ny)
    ours = create_laplacian_matrix(nx, ny)
    assert np.array_equal(ours.toarray(), theirs.toarray())
    x = np.random.default_rng(0).uniform(size=(nx * ny, 3))
    assert np.allclose(ref.measures.trace_xtLx(ours, x), ref.measures.trace_xtLx(theirs, x))   # the reference's own op


@needs_ref
def test_estimator_is_an_espm_nmf_estimator():
    """The hyperspy signal class of the reference only accepts ``NMFEstimator`` instances as learning results
    (eds_spim.py:607, 639): the drop-in registers itself as a virtual subclass once espm is imported."""
    ref = ref_import.load_reference()
    import espm_b200
    from espm_b200.estimators import register_with_espm
    assert register_with_espm()
    est = espm_b200.SmoothNMF(n_components=3, verbose=0)
    assert isinstance(est, ref.estimators.NMFEstimator)
    assert issubclass(espm_b200.SmoothNMF, ref.estimators.NMFEstimator)
    # the real subclasses are unaffected
    assert isinstance(ref.SmoothNMF(n_components=3, verbose=0), ref.estimators.NMFEstimator)


def test_get_losses_layout_with_ground_truth():
    """base.py:479-517: with true_D / true_H the record has ang_p*, mse_p* and true_KL_loss columns."""
    import espm_b200
    k = 3
    est = espm_b200.SmoothNMF(n_components=k, verbose=0, true_D=np.ones((5, k)), true_H=np.ones((k, 4)))
    est.losses_ = [1.0, 0.5]
    est.detailed_losses_ = [[0.9, 0.05, 0.05, 8.0], [0.4, 0.05, 0.05, 8.0]]
    est.rel_ = [[0.1, 0.2], [0.01, 0.02]]
    est.angles_ = [(1.0, 2.0, 3.0), (0.5, 1.0, 1.5)]
    est.mse_ = [(0.1, 0.2, 0.3), (0.01, 0.02, 0.03)]
    est.true_losses_ = [2.0, 1.5]
    arr = est.get_losses()
    assert arr.dtype.names == ("full_loss", "KL_div_loss", "log_reg_loss", "Lapl_reg_loss", "gamma", "rel_W", "rel_H",
                               "ang_p0", "ang_p1", "ang_p2", "mse_p0", "mse_p1", "mse_p2", "true_KL_loss")
    assert arr["ang_p1"].tolist() == [2.0, 1.0] and arr["true_KL_loss"].tolist() == [2.0, 1.5]
    plain = espm_b200.SmoothNMF(n_components=k, verbose=0)
    plain.losses_, plain.detailed_losses_, plain.rel_ = est.losses_, est.detailed_losses_, est.rel_
    assert plain.get_losses().dtype.names == arr.dtype.names[:7]
