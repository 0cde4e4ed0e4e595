"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol declared in
include/ica_b200.h, the host-side scalar algebra (same source as the device epilogue) matches the
reference goldens, the pyramid's banded operator matches scipy, configuration handling, error
behaviour without a GPU, and the multi-GPU sharding logic under gloo (world size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nat():
    from inverse_compositional_algorithm_b200 import _native
    _native.lib()
    return _native


def test_library_exports_every_declared_symbol(nat):
    header = open(os.path.join(ROOT, "include", "ica_b200.h")).read()
    declared = set(re.findall(r"ICA_API\s+[\w\s\*]+?\b(ica_\w+)\s*\(", header))
    assert len(declared) >= 30
    handle = ctypes.CDLL(nat._build.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/ica_b200.h but not exported"
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)


def test_constants_agree_with_native(nat):
    from inverse_compositional_algorithm_b200 import constants as cts
    c = nat.constants()
    assert (c[0], c[1], c[2], c[3]) == (cts.MAX_ITER, cts.LAMBDA_0, cts.LAMBDA_N, cts.LAMBDA_RATIO)
    assert c[4] == 12


def test_no_gpu_fails_loudly(nat):
    if nat.device_count() > 0:
        pytest.skip("a GPU is visible")
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import (
        pyramidal_inverse_compositional_algorithm)
    img = np.zeros((16, 16, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        pyramidal_inverse_compositional_algorithm(img, img, np.zeros(2), 1, 2, 0.5, 1e-3, 0, 0.0, True, 2, False)


def test_enums_match_reference():
    from inverse_compositional_algorithm_b200 import TransformType, RobustErrorFunctionType
    assert [(t.name, t.value, t.nparams()) for t in TransformType] == [
        ("TRANSLATION", 1, 2), ("EUCLIDEAN", 2, 3), ("SIMILARITY", 3, 4), ("AFFINITY", 4, 6), ("HOMOGRAPHY", 5, 8)]
    assert [(r.name, r.value) for r in RobustErrorFunctionType] == [
        ("QUADRATIC", 0), ("TRUNCATED_QUADRATIC", 1), ("GERMAN_MCCLURE", 2), ("LORENTZIAN", 3), ("CHARBONNIER", 4)]


def test_scalar_algebra_matches_reference_goldens(nat, reference_runs):
    """update_transform / zoom_in_parameters / params2matrix run on the host from the same source as the
    device epilogue (csrc/ica_transform.cuh) and must reproduce the unmodified reference's outputs."""
    from inverse_compositional_algorithm_b200 import transformation as tr, zoom as zm
    g = reference_runs
    for t in tr.TransformType:
        for p, dp, want, zwant in zip(g[f"upd/{t.name}/p"], g[f"upd/{t.name}/dp"], g[f"upd/{t.name}/out"],
                                      g[f"zoomin/{t.name}/out"]):
            q = p.copy()
            r = tr.update_transform(q, dp, t)
            assert r is q  # in place, like the reference
            np.testing.assert_allclose(q, want, rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(zm.zoom_in_parameters(p, t, 97.0, 49.0, 194.0, 97.0), zwant, rtol=1e-15)
        np.testing.assert_allclose(tr.params2matrix(g[f"upd/{t.name}/p"][0], t), g[f"p2m/{t.name}"], rtol=1e-15)
        np.testing.assert_allclose(tr.matrix2params(g[f"p2m/{t.name}"], t), g[f"upd/{t.name}/p"][0], rtol=1e-12,
                                   atol=1e-15)
    assert zm.zoom_size(97, 49, 0.5) == (48, 24)  # round-half-to-even (zoom.py:20-21)
    assert zm.zoom_size(584, 388, 0.5) == (292, 194)
    with pytest.raises(ValueError):
        zm.zoom_in_parameters(np.zeros(2), 9, 10, 10, 20, 20)
    with pytest.raises(ValueError):
        tr.update_transform(np.zeros(2), np.zeros(2), 7)


def test_jacobian_known_answers():
    """The reference's own unit tests (test/test_derivatives.py:13-68) against the package's jacobian."""
    from inverse_compositional_algorithm_b200 import derivatives as de
    from inverse_compositional_algorithm_b200 import TransformType as T
    np.testing.assert_array_equal(de.jacobian(T.TRANSLATION, 2, 2)[1, 1], [1, 0, 0, 1])
    np.testing.assert_array_equal(de.jacobian(T.EUCLIDEAN, 2, 2)[1, 1], [1, 0, -1, 0, 1, 1])
    np.testing.assert_array_equal(de.jacobian(T.SIMILARITY, 2, 2)[1, 0], [1, 0, 0, -1, 0, 1, 1, 0])
    np.testing.assert_array_equal(de.jacobian(T.AFFINITY, 2, 2)[1, 1], [1, 0, 1, 1, 0, 0, 0, 1, 0, 0, 1, 1])
    from oracle import ica_oracle as orc
    for t in T:
        np.testing.assert_array_equal(de.jacobian(t, 7, 5), orc.jacobian(t.value, 7, 5))


def test_inverse_hessian_matches_lapack_and_singular_case(nat):
    rng = np.random.default_rng(1)
    for n in (2, 3, 4, 6, 8):
        A = rng.normal(size=(n, n))
        H = A @ A.T + np.diag(10.0 ** rng.uniform(0, 8, n))
        np.testing.assert_allclose(nat.inverse_hessian(H), np.linalg.inv(H), rtol=1e-9)
    assert np.all(nat.inverse_hessian(np.zeros((4, 4))) == 0.0)  # LinAlgError branch: zero matrix


@pytest.mark.parametrize("n_in,n_out", [(64, 32), (97, 48), (388, 194), (53, 26), (1024, 512), (90, 63), (5, 2)])
def test_pyramid_operator_matches_scipy(nat, n_in, n_out):
    """The banded operator the CUDA pyramid applies per axis vs scipy's gaussian_filter1d + spline zoom
    (what skimage.transform.rescale does); 1e-5 relative is the north-star budget, the band is cut at 1e-9."""
    from scipy import ndimage as ndi
    A, taps, fast = nat.resample_operator(n_in, n_out)
    rng = np.random.default_rng(n_in)
    x = rng.uniform(0, 255, (n_in, 3))
    f = n_in / n_out
    sigma = max(0.0, (f - 1) / 2)
    g = ndi.gaussian_filter1d(x, sigma, axis=0, mode="constant", cval=0) if sigma > 1e-15 else x
    want = ndi.zoom(g, (1 / f, 1), order=3, mode="grid-constant", cval=0, grid_mode=True)
    got = A @ x
    assert got.shape == want.shape
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-6
    assert taps <= 64
    if n_in == 2 * n_out and n_out >= 32:
        assert fast[1] - fast[0] >= n_out - 24  # exact 2:1 levels: uniform weights away from the borders
        lo, hi, s0 = fast
        np.testing.assert_allclose(A[lo + 3, 2 * (lo + 3) + s0:2 * (lo + 3) + s0 + taps],
                                   A[hi - 2, 2 * (hi - 2) + s0:2 * (hi - 2) + s0 + taps], atol=1e-8)


@pytest.mark.parametrize("n_in,n_out", [(600, 24), (1000, 30), (1500, 40)])
def test_resample_operator_strong_downscaling(nat, n_in, n_out):
    """nu around 0.04-0.06: the Gaussian radius (4 sigma = 2 (1/nu - 1)) exceeds the 48 columns the band builder used to
    keep: the band is now sized from the radius and the operator still matches scipy."""
    from scipy import ndimage as ndi
    A, taps, _ = nat.resample_operator(n_in, n_out)
    x = np.random.default_rng(n_in).uniform(0, 255, (n_in, 2))
    f = n_in / n_out
    g = ndi.gaussian_filter1d(x, (f - 1) / 2, axis=0, mode="constant", cval=0)
    want = ndi.zoom(g, (1 / f, 1), order=3, mode="grid-constant", cval=0, grid_mode=True)
    assert np.abs(A @ x - want).max() / np.abs(want).max() < 2e-6
    if n_in / n_out > 26:
        assert taps > 97      # wider than the old fixed band of 2 * 48 + 1 columns


def test_configuration_handler_roundtrip(tmp_path):
    from inverse_compositional_algorithm_b200 import configuration_handler as cfh
    from inverse_compositional_algorithm_b200 import TransformType, RobustErrorFunctionType
    path = tmp_path / "config.ini"
    cfh.create_config_file(str(path))
    cfg = cfh.read_config_file(str(path))
    assert set(cfg) == {"inverse_compositional_algorithm", "robust_inverse_compositional_algorithm",
                        "pyramidal_inverse_compositional_algorithm"}
    pica = cfg["pyramidal_inverse_compositional_algorithm"]
    assert pica["TOL"] == 1e-3 and pica["transform_type"] == TransformType.EUCLIDEAN
    assert pica["pyramid_levels"] == 2 and pica["nu"] == 0.5 and pica["robust_type"] == RobustErrorFunctionType.QUADRATIC
    assert cfg["robust_inverse_compositional_algorithm"]["robust_type"] == RobustErrorFunctionType.CHARBONNIER
    # the reference's own root config.ini carries inline comments; they are tolerated here
    path.write_text("[InverseCompositionalAlgorithm]\ntol = 1e-3\ntransform_type = EUCLIDEAN #TRANSLATION, ...\n"
                    "verbose = True\n[RobustInverseCompositionalAlgorithm]\ntol = 1e-3\ntransform_type = AFFINITY\n"
                    "robust_type = LORENTZIAN #x\nlambda = 0.0\nverbose = True\n"
                    "[PyramidalInverseCompositionalAlgorithm]\ntol = 1e-3\ntransform_type = HOMOGRAPHY\n"
                    "pyramid_levels = 5\nnu = 0.5\nrobust_type = QUADRATIC\nlambda = 0.0\nverbose = True\n")
    cfg = cfh.read_config_file(str(path))
    assert cfg["inverse_compositional_algorithm"]["transform_type"] == TransformType.EUCLIDEAN
    assert cfg["pyramidal_inverse_compositional_algorithm"]["pyramid_levels"] == 5


def test_shard_range_partitions():
    from inverse_compositional_algorithm_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 4096, 1025):
        for w in (1, 2, 3, 8):
            blocks = [shard_range(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_row_band_partitions():
    """Row-sharded mode: the bands of tile rows of the ranks tile [0, tiles_y) without overlap (some may be
    empty when there are fewer tile rows than ranks)."""
    from inverse_compositional_algorithm_b200.sharding import row_band
    for tiles_y in (0, 1, 5, 74, 586):
        for w in (1, 2, 3, 8):
            bands = [row_band(tiles_y, r, w) for r in range(w)]
            assert bands[0][0] == 0 and bands[-1][1] == tiles_y
            assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
            sizes = [b - a for a, b in bands]
            assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        row_band(4, 2, 2)


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from inverse_compositional_algorithm_b200.sharding import register_sharded, shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
n = 7
I1 = np.arange(n * 4, dtype=np.float32).reshape(n, 2, 2, 1)
def fake_register(a, b, types):            # stands in for the GPU plan: results are a function of the pair only
    k = a.reshape(len(a), -1).sum(1)
    p = np.zeros((len(a), 8)); p[:, 0] = k; p[:, 1] = [t for t in types]
    return p, k * 0.5, np.stack([k.astype(np.int32), k.astype(np.int32) + 1], 1)
p, err, iters = register_sharded(I1, I1, list(range(n)), fake_register)
want = I1.reshape(n, -1).sum(1)
assert p.shape == (n, 8) and np.array_equal(p[:, 0], want) and np.array_equal(p[:, 1], np.arange(n))
assert np.array_equal(err, want * 0.5) and np.array_equal(iters[:, 1], want.astype(np.int32) + 1)
lo, hi = shard_range(n, dist.get_rank(), 2)
assert (lo, hi) == ((0, 4) if dist.get_rank() == 0 else (4, 7))
dist.barrier(); dist.destroy_process_group()
print("ok")
'''


def test_sharded_registration_gloo_world2(tmp_path):
    """N > 1 path on CPU: two gloo ranks each register their block of pairs, results are all-gathered."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o


def test_zoom_out_operator_reproduces_scipy_chain():
    """Host logic of zoom.zoom_out: the banded 1-D operator (built in C++: build_zoom_out_1d) applied along
    both axes equals the direct Gaussian + cubic-spline resampling of the oracle."""
    from oracle import ica_oracle as orc
    from inverse_compositional_algorithm_b200.zoom import zoom_out_operator
    rng = np.random.default_rng(11)
    img = rng.uniform(0, 255, (75, 98, 2))
    for f in (0.5, 0.6):
        ys, yw = zoom_out_operator(75, f)
        xs, xw = zoom_out_operator(98, f)
        assert yw.shape[1] <= 128 and xw.shape[1] <= 128
        ay = np.zeros((ys.size, 75)); ax = np.zeros((xs.size, 98))
        for o in range(ys.size): ay[o, ys[o]:ys[o] + yw.shape[1]] = yw[o]
        for o in range(xs.size): ax[o, xs[o]:xs[o] + xw.shape[1]] = xw[o]
        got = np.einsum("oy,yxc,px->opc", ay, img, ax)
        want = orc.zoom_out(img, f)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


def test_ipol_point_interpolation_matches_reference_golden():
    """Scalar helper ``bicubic_interpolation_point`` (host logic) against the reference's whole-image outputs."""
    from inverse_compositional_algorithm_b200 import bicubic_interpolation as bi
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ipol_warp.npz")))
    img, p, want = g["image"], g["params_0"], g["out_0"]          # translation, delta 2, NaN outside
    ny, nx, nz = img.shape
    for (i, j) in [(5, 7), (10, 20), (15, 3)]:
        x, y = j + p[0], i + p[1]
        for k in range(nz):
            assert abs(bi.bicubic_interpolation_point(img, x, y, nx, ny, nz, k) - want[i, j, k]) <= 1e-10
