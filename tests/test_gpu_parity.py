"""GPU parity tests: the CUDA path (through the C-ABI / the Python drop-in) against the CPU
oracle on the same inputs and against the committed golden fixtures.

Tolerances (BASELINE.json north_star): pyramid and gradients within 1e-5 relative
(max|a-b| / max|b|); final motion parameters within an end-point error of 1e-3 px over the
image domain.  Images are float32 on the device, parameters and reductions float64.
"""
import numpy as np
import pytest

from oracle import ica_oracle as orc

pytestmark = pytest.mark.gpu

REL_1E5 = 1e-5
EPE_TOL = 1e-3
# per-iteration |dp| against the float64 oracle / reference.  The device stores images (and pyramid levels) in float32
# -- the north star's 1e-5 budget -- so an iterate differs from the float64 one by ~1e-6 px, and |dp|, the size of the
# NEXT step, inherits that absolutely: measured worst cases over the suite (B200, round 2) are 1.7e-3 relative on a
# |dp| of 5e-3 and 1.5e-5 absolute on a |dp| of 1.4e-2, with identical iteration counts and final EPE <= 2.2e-6 px.
DP_RTOL, DP_ATOL = 2.5e-3, 2e-6
HB8192_DP_RTOL = 1e-4      # one solve of the ill-conditioned 8 x 8 system at 8192^2 (measured 5e-6)
NOTEBOOK_ERR_RTOL = 2.5e-3   # the notebooks' stored |Dp| lines (was 0.05 in round 1)


@pytest.fixture(scope="module")
def nat():
    from inverse_compositional_algorithm_b200 import _native
    _native.require_gpu()
    return _native


def rel_err(a, b):
    return float(np.nanmax(np.abs(a - b)) / np.nanmax(np.abs(b)))


# ------------------------------------------------------------------ K2-lite: the warp alone
def test_warp_matches_reference_goldens(nat, reference_runs):
    from inverse_compositional_algorithm_b200.bicubic_interpolation import bicubic_interpolation_skimage
    from inverse_compositional_algorithm_b200.transformation import TransformType
    g = reference_runs
    keys = [k for k in g if k.startswith("warp/") and k.endswith("/out")]
    assert len(keys) == 6
    for key in keys:
        t = TransformType[key.split("/")[1]]
        got = bicubic_interpolation_skimage(g["warp/img"], g[key[:-4] + "/p"], t, True, 5)
        want = g[key]
        assert np.array_equal(np.isnan(got), np.isnan(want)), key  # NaN footprint, SURVEY Q2
        assert rel_err(got, want) <= REL_1E5, key


def test_warp_rubber_whale_subpixel(nat, rubber_whale):
    img = rubber_whale["rubber_whale"].astype(np.float64)
    for ttype, p in ((orc.TRANSLATION, [-9.99, -5.3]), (orc.EUCLIDEAN, [3.2, -1.7, -0.1]),
                     (orc.SIMILARITY, [0.5, 0.25, 0.11, 0.002]),
                     (orc.HOMOGRAPHY, [0.01, 0.02, 2.2, -0.015, 0.012, -1.1, 2e-5, -1e-5])):
        want = orc.warp_bicubic(img, np.array(p), ttype)
        got = nat.warp(img, orc.params2matrix(p, ttype)).astype(np.float64)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert rel_err(got, want) <= REL_1E5


# ------------------------------------------------------------------ K0: one pyramid level
@pytest.mark.parametrize("key", ["37x53", "64x48", "97x131"])
def test_rescale_matches_reference_goldens(nat, reference_runs, key):
    from inverse_compositional_algorithm_b200.zoom import rescale
    got = rescale(reference_runs[f"rescale/{key}/in"], 0.5)
    want = reference_runs[f"rescale/{key}/out"]
    assert got.shape == want.shape
    assert rel_err(got, want) <= REL_1E5


def test_pyramid_cascade_rubber_whale(nat, rubber_whale):
    img = rubber_whale["rubber_whale_rt"].astype(np.float64)
    levels = orc.build_pyramid(img, 3, 0.5)
    cur = img
    for s in (1, 2):
        cur = nat.rescale(cur, 0.5).astype(np.float64)
        assert cur.shape == levels[s].shape
        assert rel_err(cur, levels[s]) <= REL_1E5


def test_rescale_other_factor_and_gray(nat):
    rng = np.random.default_rng(5)
    img = rng.uniform(0, 255, (90, 70, 1)).astype(np.float32)
    want = orc.sk.rescale(np.repeat(img, 3, 2).astype(np.float64), 0.7)[:, :, :1]
    got = nat.rescale(img, 0.7)
    assert got.shape == want.shape and rel_err(got, want) <= REL_1E5


# ------------------------------------------------------------------ gradient + frame
def test_gradient_and_frame(nat, rubber_whale):
    img = rubber_whale["rubber_whale_tr"].astype(np.float64)
    for nanif, delta in ((True, 10), (False, 10), (True, 0)):
        Ix, Iy = orc.gradient_with_frame(img, nanif, delta)
        gx, gy = nat.gradient(img, delta, nanif)
        assert np.array_equal(np.isnan(gx), np.isnan(Ix)) and np.array_equal(np.isnan(gy), np.isnan(Iy))
        assert rel_err(gx, Ix) <= REL_1E5 and rel_err(gy, Iy) <= REL_1E5


# ------------------------------------------------------------------ one H / b evaluation
@pytest.mark.parametrize("ttype,rtype", [(orc.TRANSLATION, orc.CHARBONNIER), (orc.EUCLIDEAN, orc.LORENTZIAN),
                                         (orc.SIMILARITY, orc.GERMAN_MCCLURE), (orc.AFFINITY, orc.QUADRATIC),
                                         (orc.HOMOGRAPHY, orc.LORENTZIAN), (orc.HOMOGRAPHY, orc.TRUNCATED_QUADRATIC)])
def test_hessian_and_b_one_iteration(nat, ttype, rtype):
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.transformation import TransformType
    I1, I2, p_gt = synthetic.make_pair(21, 100, 140, 3, TransformType(ttype), max_shift=3.0, margin=32)
    p = 0.6 * p_gt
    lam = 20.0
    H, b = nat.hessian_b(I1, I2, ttype, p, rtype, lam, 7, True)
    I1d, I2d = I1.astype(np.float64), I2.astype(np.float64)
    n = orc.nparams(ttype)
    Ix, Iy = orc.gradient_with_frame(I1d, True, 7)
    DIJ = orc.steepest_descent_images(Ix, Iy, orc.jacobian(ttype, 140, 100), n)
    DI = orc.warp_bicubic(I2d, p, ttype) - I1d
    rho = orc.robust_error_function(DI, lam, rtype)
    Hw = orc.hessian_robust(DIJ, rho)
    bw = orc.independent_vector_robust(DIJ, DI, rho)
    # entries of H span many orders of magnitude: compare each one relative to its own scale
    scale = np.sqrt(np.outer(np.diag(Hw), np.diag(Hw)))
    assert np.max(np.abs(H - Hw) / scale) <= 1e-5
    dpw = np.linalg.solve(Hw, bw)
    dpg = np.linalg.solve(H, b)
    np.testing.assert_allclose(dpg, dpw, rtol=1e-3, atol=1e-6 * np.abs(dpw).max())


# ------------------------------------------------------------------ whole registrations
def _epe(pa, pb, ttype, nx, ny):
    return orc.end_point_error(pa, pb, ttype, nx, ny)[1]


@pytest.mark.parametrize("idx", range(14))
def test_registration_matches_reference_runs(nat, reference_runs, idx):
    """The unmodified reference's results (tests/golden/reference_runs.npz) on seeded synthetic
    pairs, every transform type and error function; the oracle is re-run for the trajectory."""
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import (
        pyramidal_inverse_compositional_algorithm)
    from inverse_compositional_algorithm_b200.image_optimisation import RobustErrorFunctionType
    from inverse_compositional_algorithm_b200.transformation import TransformType
    g = reference_runs
    name = str(g["names"][idx])
    seed, H, W, tt, rt, nscales, lam, occ, max_shift, delta = g[name + "/cfg"]
    t = TransformType(int(tt))
    I1, I2, p_gt = synthetic.make_pair(int(seed), int(H), int(W), 3, t, max_shift=float(max_shift),
                                       occlusion=float(occ), margin=32)
    p, err, DI, Iw = pyramidal_inverse_compositional_algorithm(
        I1, I2, np.zeros(t.nparams()), t, int(nscales), 0.5, 1e-3, RobustErrorFunctionType(int(rt)),
        float(lam), True, int(delta), False)
    want = g[name + "/p"]
    epe = _epe(p, want, t.value, int(W), int(H))
    assert epe <= EPE_TOL, (name, epe, p, want)
    assert DI.shape == I1.shape and DI.dtype == np.float64
    # the reference's returned Iw: same NaN footprint, same values at the probe points
    yy = np.linspace(0, int(H) - 1, 9).astype(int)
    xx = np.linspace(0, int(W) - 1, 11).astype(int)
    probe = Iw[np.ix_(yy, xx)]
    wantp = g[name + "/Iw_probe"]
    assert np.array_equal(np.isnan(probe), np.isnan(wantp))
    assert np.nanmax(np.abs(probe - wantp)) <= 0.05  # grey levels; p differs by <= 1e-3 px


@pytest.mark.parametrize("idx", [0, 4, 9, 10, 13])
def test_trajectory_matches_oracle(nat, reference_runs, idx):
    """Per-iteration |dp| and p against the golden trajectory (same iteration count, each p within
    the EPE budget)."""
    from inverse_compositional_algorithm_b200 import _native, synthetic
    from inverse_compositional_algorithm_b200.transformation import TransformType
    g = reference_runs
    name = str(g["names"][idx])
    seed, H, W, tt, rt, nscales, lam, occ, max_shift, delta = g[name + "/cfg"]
    t = TransformType(int(tt))
    I1, I2, _ = synthetic.make_pair(int(seed), int(H), int(W), 3, t, max_shift=float(max_shift),
                                    occlusion=float(occ), margin=32)
    plan = _native.Plan(batch=1, height=int(H), width=int(W), channels=3, nscales=int(nscales), nu=0.5,
                        transform_type=t.value, robust_type=int(rt), robust_loop=int(rt) != 0,
                        lambda_=float(lam), tol=1e-3, max_iter=30, delta=int(delta), nanifoutside=True,
                        record_trajectory=True)
    plan.run_host(I1[None], I2[None])
    traj = plan.trajectory()[0]
    want = g[name + "/traj"]
    n = t.nparams()
    # last |dp| of each scale decides the stopping rule; report the margin if counts differ
    assert len(traj) == len(want), (len(traj), len(want), want[:, 1])
    nx, ny = plan.level_shapes()
    for row, w in zip(traj, want):
        assert int(row[0]) == int(w[0])
        s = int(w[0])
        assert _epe(row[4:4 + n], w[3:3 + n], t.value, int(nx[s]), int(ny[s])) <= EPE_TOL
        if not np.isnan(w[2]):
            np.testing.assert_allclose(row[3], w[2], rtol=1e-12)  # lambda schedule
    # per-iteration |dp| (float32 images on the device, float64 in the reference: the budget is their rounding)
    np.testing.assert_allclose(traj[:, 2], want[:, 1], rtol=DP_RTOL, atol=DP_ATOL)
    plan.close()


SAMPLES = {"rubber_whale_tr": (orc.TRANSLATION, "tr"), "rubber_whale_rt": (orc.EUCLIDEAN, "rt"),
           "rubber_whale_eu": (orc.EUCLIDEAN, "eu"), "rubber_whale_zo": (orc.SIMILARITY, "zo")}


@pytest.mark.parametrize("sample", list(SAMPLES))
def test_notebook_pyramidal_charbonnier(nat, rubber_whale, notebook_runs, sample):
    """The reference authors' stored run (robust.ipynb cell 15): uint8 images from disk, 3 scales,
    CHARBONNIER.  Same number of iterations per scale and final p within 1e-3 px."""
    from inverse_compositional_algorithm_b200 import _native
    entries = [r for r in notebook_runs["inverse_compositional_algorithm_robust.ipynb"]
               if r["cell"] == 15 and r["sample"] == sample][0]["entries"]
    ttype, suffix = SAMPLES[sample]
    I1 = rubber_whale["rubber_whale_" + suffix]
    I2 = rubber_whale["rubber_whale"]
    plan = _native.Plan(batch=1, height=388, width=584, channels=3, nscales=3, nu=0.5,
                        transform_type=ttype, robust_type=orc.CHARBONNIER, robust_loop=True, lambda_=0.0,
                        tol=1e-3, max_iter=30, delta=10, nanifoutside=True, record_trajectory=True)
    p, err, iters, _, _ = plan.run_host(I1[None], I2[None])  # uint8 straight through the ABI
    traj = plan.trajectory()[0]
    n = orc.nparams(ttype)
    assert len(traj) == len(entries)
    assert _epe(p[0, :n], entries[-1]["p"], ttype, 584, 388) <= EPE_TOL
    np.testing.assert_allclose(err[0], entries[-1]["err"], rtol=NOTEBOOK_ERR_RTOL)
    # every stored line of the notebook: |Dp| and p, iteration by iteration
    for row, e in zip(traj, entries):
        np.testing.assert_allclose(row[2], e["err"], rtol=NOTEBOOK_ERR_RTOL, atol=DP_ATOL)
    plan.close()


def test_notebook_single_scale_robust_translation(nat, rubber_whale, notebook_runs):
    """robust.ipynb cell 13 through the drop-in function (18 iterations, in-place p)."""
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import (
        robust_inverse_compositional_algorithm)
    from inverse_compositional_algorithm_b200.image_optimisation import RobustErrorFunctionType
    from inverse_compositional_algorithm_b200.transformation import TransformType
    entries = [r for r in notebook_runs["inverse_compositional_algorithm_robust.ipynb"]
               if r["cell"] == 13][0]["entries"]
    p = np.zeros(2)
    pr, err, DI, Iw = robust_inverse_compositional_algorithm(
        I1=rubber_whale["rubber_whale_tr"], I2=rubber_whale["rubber_whale"], p=p,
        transform_type=TransformType.TRANSLATION, TOL=1e-3, robust_type=RobustErrorFunctionType.CHARBONNIER,
        lambda_=0.0, nanifoutside=True, delta=10, verbose=False)
    assert pr is p  # the single-scale functions mutate their argument
    assert _epe(p, entries[-1]["p"], orc.TRANSLATION, 584, 388) <= EPE_TOL
    np.testing.assert_allclose(err, entries[-1]["err"], rtol=NOTEBOOK_ERR_RTOL)


def test_quadratic_single_scale_matches_oracle(nat, rubber_whale):
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import inverse_compositional_algorithm
    from inverse_compositional_algorithm_b200.transformation import TransformType
    I2 = rubber_whale["rubber_whale"][100:260, 200:420].astype(np.float64)
    I1 = orc.transform_image(I2, orc.AFFINITY, [1.5, -0.8, 0.01, 0.005, -0.004, 0.008])
    want, werr, wDI, wIw = orc.ica_quadratic(I1, I2, np.zeros(6), orc.AFFINITY, 1e-3, True, 10)
    p, err, DI, Iw = inverse_compositional_algorithm(I1, I2, np.zeros(6), TransformType.AFFINITY, 1e-3,
                                                     True, 10, False)
    assert _epe(p, want, orc.AFFINITY, 220, 160) <= EPE_TOL
    assert np.array_equal(np.isnan(Iw), np.isnan(wIw))
    assert np.nanmax(np.abs(Iw - wIw)) <= 0.05 and np.nanmax(np.abs(DI - wDI)) <= 0.05


# ------------------------------------------------------------------ batches, gray, edge cases
def test_mixed_batch_equals_single_runs(nat):
    """Ragged batch (similarity/affinity mix, BASELINE config 3 in miniature, gray images): each
    pair gets the result of its own single run."""
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.transformation import TransformType
    types = [TransformType.SIMILARITY, TransformType.AFFINITY] * 3
    pairs = [synthetic.make_pair(40 + i, 120, 160, 1, t, max_shift=4.0, margin=32) for i, t in enumerate(types)]
    I1 = np.stack([a for a, _, _ in pairs])
    I2 = np.stack([b for _, b, _ in pairs])
    p, err, iters = register_batch(I1, I2, types, nscales=3, delta=5)
    for i, t in enumerate(types):
        ps, es, its = register_batch(I1[i:i + 1], I2[i:i + 1], t, nscales=3, delta=5)
        assert _epe(p[i], ps[0], t.value, 160, 120) <= 1e-6
        assert np.array_equal(iters[i], its[0])
        # and the oracle on the gray image replicated to RGB
        po, _, _, _ = orc.ica_pyramidal(np.repeat(I1[i], 3, 2), np.repeat(I2[i], 3, 2), np.zeros(t.nparams()),
                                        t.value, 3, 0.5, 1e-3, orc.QUADRATIC, 0.0, True, 5)
        assert _epe(p[i, :t.nparams()], po, t.value, 160, 120) <= EPE_TOL
        assert _epe(p[i, :t.nparams()], pairs[i][2], t.value, 160, 120) <= 0.05  # ground truth


def test_gray_as_rgb_robust_weights(nat):
    """Gray input must behave as its x3 replication also for robust functions (rho' depends on
    the channel sum)."""
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType.HOMOGRAPHY
    I1, I2, _ = synthetic.make_pair(77, 128, 128, 1, t, max_shift=3.0, margin=32)
    pg, eg, ig = register_batch(I1[None], I2[None], t, nscales=2, robust_type=3, delta=5)
    pc, ec, ic = register_batch(np.repeat(I1, 3, 2)[None], np.repeat(I2, 3, 2)[None], t, nscales=2,
                                robust_type=3, delta=5)
    assert np.array_equal(ig, ic)
    assert _epe(pg[0], pc[0], t.value, 128, 128) <= 1e-5


def test_deterministic(nat):
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType.HOMOGRAPHY
    pairs = [synthetic.make_pair(90 + i, 96, 128, 3, t, max_shift=3.0, margin=32) for i in range(4)]
    I1 = np.stack([a for a, _, _ in pairs]); I2 = np.stack([b for _, b, _ in pairs])
    a = register_batch(I1, I2, t, nscales=3, robust_type=2, delta=5)
    b = register_batch(I1, I2, t, nscales=3, robust_type=2, delta=5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_singular_hessian_and_tiny_images(nat):
    """Constant images: H is exactly singular -> zero inverse -> dp = 0, error = 0, loop exits after
    one iteration with p unchanged (de.py:125-129).  Also shapes smaller than a tile / the frame."""
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import (
        pyramidal_inverse_compositional_algorithm, register_batch)
    from inverse_compositional_algorithm_b200.transformation import TransformType
    flat = np.full((40, 50, 3), 17.0)
    p, err, DI, Iw = pyramidal_inverse_compositional_algorithm(flat, flat, np.zeros(6), TransformType.AFFINITY,
                                                               2, 0.5, 1e-3, 4, 0.0, True, 3, False)
    assert err == 0.0 and np.all(p == 0.0)
    rng = np.random.default_rng(3)
    tiny = rng.uniform(0, 255, (1, 9, 11, 3)).astype(np.float32)
    p, err, iters = register_batch(tiny, tiny, TransformType.TRANSLATION, nscales=1, delta=2)
    po, eo, _, _ = orc.ica_quadratic(tiny[0], tiny[0], np.zeros(2), orc.TRANSLATION, 1e-3, True, 2)
    np.testing.assert_allclose(p[0, :2], po, atol=1e-4)


def test_errors_like_the_reference(nat):
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import (
        inverse_compositional_algorithm, pyramidal_inverse_compositional_algorithm)
    from inverse_compositional_algorithm_b200.transformation import TransformType
    a = np.zeros((20, 30, 3)); b = np.zeros((20, 31, 3)); g = np.zeros((20, 30))
    with pytest.raises(ValueError):
        inverse_compositional_algorithm(a, b, np.zeros(2), TransformType.TRANSLATION, 1e-3, True, 2, False)
    with pytest.raises(ValueError):
        inverse_compositional_algorithm(g, g, np.zeros(2), TransformType.TRANSLATION, 1e-3, True, 2, False)
    with pytest.raises(ValueError):
        pyramidal_inverse_compositional_algorithm(a, a, np.zeros(2), TransformType.TRANSLATION, 2, 0.5, 0.01,
                                                  0, 0.0, True, 2, False)


def test_full_size_c2_properties(nat):
    """BASELINE config 2 shape (1024x1024 RGB, HOMOGRAPHY, LORENTZIAN, 5 scales): too slow for the
    oracle inside a unit test, so check size-independent properties: recovers the ground-truth
    motion, is bit-reproducible, and registering the pair against itself returns identity."""
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType.HOMOGRAPHY
    I1, I2, p_gt = synthetic.make_pair(1234, 1024, 1024, 3, t, margin=64)
    a = register_batch(I1[None], I2[None], t, nscales=5, robust_type=3, delta=10)
    b = register_batch(I1[None], I2[None], t, nscales=5, robust_type=3, delta=10)
    assert np.array_equal(a[0], b[0])
    assert _epe(a[0][0], p_gt, t.value, 1024, 1024) <= 0.05
    c = register_batch(I2[None], I2[None], t, nscales=5, robust_type=3, delta=10)
    assert _epe(c[0][0], np.zeros(8), t.value, 1024, 1024) <= 1e-6


# ------------------------------------------------------------------ row-sharded single pair (BASELINE config 5)
@pytest.mark.parametrize("shape,channels,world,rtype", [((200, 256), 3, 3, 3), ((96, 128), 1, 8, 0),
                                                        ((300, 200), 1, 2, 2)])
def test_row_sharded_equals_unsharded(nat, shape, channels, world, rtype):
    """One pair split by bands of rows over `world` ranks (emulated one after the other on this GPU, same
    kernels and band arithmetic as the NCCL path): the moments summed over bands equal the unsharded sums up
    to fp64 regrouping, so iteration counts are identical and the motion agrees to far below the parity bar.  world = 8 on a
    small image leaves some ranks without tiles at the coarse levels."""
    import torch
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.sharding import register_row_sharded
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType.HOMOGRAPHY
    h, w = shape
    I1, I2, p_gt = synthetic.make_pair(500 + world, h, w, channels, t, max_shift=3.0, margin=32)
    ref_p, ref_err, ref_it = register_batch(I1[None], I2[None], t, nscales=3, robust_type=rtype, delta=5)
    stats = {}
    p, err, iters = register_row_sharded(torch.from_numpy(I1).cuda(), torch.from_numpy(I2).cuda(), t, nscales=3,
                                         robust_type=rtype, delta=5, emulate_ranks=world, stats=stats)
    assert np.array_equal(iters, ref_it[0])
    assert stats["iterations"] == int(ref_it[0].sum())
    epe = _epe(p, ref_p[0], t.value, w, h)
    assert epe <= 1e-6, epe          # fp64 regrouping of the sums, amplified by cond(H) of a homography
    assert abs(err - ref_err[0]) <= 1e-6


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch
import torch.distributed as dist
from inverse_compositional_algorithm_b200 import synthetic
from inverse_compositional_algorithm_b200.sharding import register_row_sharded
from inverse_compositional_algorithm_b200.transformation import TransformType
rank = int(sys.argv[3])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=rank, world_size=2)
t = TransformType.HOMOGRAPHY
I1, I2, _ = synthetic.make_pair(901, 384, 512, 1, t, max_shift=3.0, margin=32)
a, b = torch.from_numpy(I1).cuda(), torch.from_numpy(I2).cuda()
p, err, iters = register_row_sharded(a, b, t, nscales=3, robust_type=3, delta=5)
# the same registration with the exchange inside the device-side loop (peer memory over NVLink, no NCCL call per iteration)
st = {}
pp, perr, piters = register_row_sharded(a, b, t, nscales=3, robust_type=3, delta=5, exchange="peer", stats=st)
np.savez(sys.argv[4] % rank, p=p, err=err, iters=iters, pp=pp, piters=piters, xus=st["exchange_us_mean"])
dist.barrier(); dist.destroy_process_group()
print("ok")
"""


def test_row_sharded_nccl_two_gpus(nat, tmp_path):
    """The real thing on two GPUs: per-iteration NCCL allreduce of the moment sums; both ranks end with
    identical parameters, equal to the single-GPU run."""
    import os, subprocess, sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.transformation import TransformType
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER)
    out = str(tmp_path / "rank%d.npz")
    port = str(31500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), root, port, str(r), out], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o
    r0, r1 = np.load(out % 0), np.load(out % 1)
    assert np.array_equal(r0["p"], r1["p"]) and np.array_equal(r0["iters"], r1["iters"])
    t = TransformType.HOMOGRAPHY
    I1, I2, _ = synthetic.make_pair(901, 384, 512, 1, t, max_shift=3.0, margin=32)
    ref_p, _, ref_it = register_batch(I1[None], I2[None], t, nscales=3, robust_type=3, delta=5)
    assert np.array_equal(r0["iters"], ref_it[0])
    assert _epe(r0["p"], ref_p[0], t.value, 512, 384) <= 1e-6
    # peer exchange: bit-identical on both ranks and equal to the NCCL path (both add the two bands' sums in rank order)
    assert np.array_equal(r0["pp"], r1["pp"]) and np.array_equal(r0["piters"], r1["piters"])
    assert np.array_equal(r0["piters"], ref_it[0])
    assert _epe(r0["pp"], ref_p[0], t.value, 512, 384) <= 1e-6
    print("peer exchange latency (us, mean):", float(r0["xus"]), float(r1["xus"]))


def test_row_sharded_peer_exchange_single_rank(nat):
    """The device-side exchange path with a group of one (the rank publishes to and waits on its own buffer): the whole
    registration is one graph launch of {K2, K3 with exchange}; result identical to the plain device-side loop."""
    import torch
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.sharding import register_row_sharded
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType.HOMOGRAPHY
    I1, I2, _ = synthetic.make_pair(902, 300, 400, 1, t, max_shift=3.0, margin=32)
    ref_p, ref_err, ref_it = register_batch(I1[None], I2[None], t, nscales=3, robust_type=3, delta=5)
    for rep in range(2):        # the second run reuses the plan, its graph and the next range of sequence numbers
        st = {}
        p, err, iters = register_row_sharded(torch.from_numpy(I1).cuda(), torch.from_numpy(I2).cuda(), t, nscales=3,
                                             robust_type=3, delta=5, exchange="peer", stats=st)
        assert np.array_equal(iters, ref_it[0]) and np.array_equal(p, ref_p[0]) and err == ref_err[0]
        assert st["launched_iterations"] == int(ref_it[0].sum())


@pytest.mark.parametrize("shape,channels,ttype_name,rtype", [((93, 121), 3, "HOMOGRAPHY", 3), ((77, 101), 1, "AFFINITY", 0),
                                                             ((64, 67), 3, "SIMILARITY", 4)])
def test_unaligned_shapes_match_oracle(nat, shape, channels, ttype_name, rtype):
    """Image rows that are not multiples of 16 bytes: the iterate kernel then works on padded level-0 copies (its TMA
    tensor maps need 16-byte row pitches) and the pyramid takes its unaligned path; results must still match the oracle."""
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType[ttype_name]
    h, w = shape
    assert (w * channels) % 4 != 0
    pairs = [synthetic.make_pair(700 + i, h, w, channels, t, max_shift=2.0, margin=32) for i in range(3)]
    I1 = np.stack([a for a, _, _ in pairs]); I2 = np.stack([b for _, b, _ in pairs])
    p, err, iters = register_batch(I1, I2, t, nscales=2, robust_type=rtype, delta=5)
    for i in range(3):
        a, b = (np.repeat(x, 3, 2) if channels == 1 else x for x in (I1[i], I2[i]))
        trace = []
        po, eo, _, _ = orc.ica_pyramidal(a.astype(np.float64), b.astype(np.float64), np.zeros(t.nparams()), t.value, 2, 0.5,
                                         1e-3, rtype, 0.0, True, 5, trace=trace)
        assert _epe(p[i, :t.nparams()], po, t.value, w, h) <= EPE_TOL
        assert int(iters[i].sum()) == len(trace)


@pytest.mark.parametrize("theta,rtype", [(0.30, 0), (-0.22, 3)])
def test_large_rotation_takes_global_path(nat, theta, rtype):
    """A rotation so strong that the window of I2 a 64x11 tile reaches does not fit the staged box: every pixel then
    samples I2 from global memory (the kernel's generic path).  Same answer as the oracle from the same start."""
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import (
        inverse_compositional_algorithm, robust_inverse_compositional_algorithm)
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType.EUCLIDEAN
    p_gt = np.array([3.0, -2.0, theta])
    I1, I2, _ = synthetic.make_pair(800, 150, 190, 3, t, margin=96, p_gt=p_gt)
    p0 = p_gt + np.array([0.6, -0.4, 0.004])
    a, b = I1.astype(np.float64), I2.astype(np.float64)
    if rtype == 0:
        p, err, DI, Iw = inverse_compositional_algorithm(a, b, p0.copy(), t, 1e-3, True, 5, False)
        po, eo, _, _ = orc.ica_quadratic(a, b, p0.copy(), t.value, 1e-3, True, 5)
    else:
        p, err, DI, Iw = robust_inverse_compositional_algorithm(a, b, p0.copy(), t, 1e-3, rtype, 0.0, True, 5, False)
        po, eo, _, _ = orc.ica_robust(a, b, p0.copy(), t.value, 1e-3, rtype, 0.0, True, 5)
    assert _epe(p, po, t.value, 190, 150) <= EPE_TOL
    assert _epe(p, p_gt, t.value, 190, 150) <= 0.1


# ------------------------------------------------------------------ helper API on materialised arrays
def test_helper_api_matches_oracle(nat, rubber_whale):
    """The building blocks callers of the reference import next to the drivers (io.rhop, io.robust_error_function,
    io.steepest_descent_images, de.hessian[_robust], io.independent_vector[_robust], tr.transform_image): same
    names and argument order, computed on the GPU in float64, compared with the oracle's restatement."""
    from inverse_compositional_algorithm_b200 import derivatives as de, image_optimisation as io, transformation as tr
    from inverse_compositional_algorithm_b200.transformation import TransformType
    img = rubber_whale["rubber_whale"][40:120, 60:170].astype(np.float64)
    rng = np.random.default_rng(5)
    ny, nx, nz = img.shape
    for t in (TransformType.AFFINITY, TransformType.HOMOGRAPHY):
        n = t.nparams()
        Ix, Iy = orc.gradient_with_frame(img, True, 4)            # NaN frame: the zero-fill rules matter
        J = de.jacobian(t, nx, ny)
        assert np.array_equal(J, orc.jacobian(t.value, nx, ny))
        DIJ = io.steepest_descent_images(Ix, Iy, J, n)
        want = orc.steepest_descent_images(Ix, Iy, J, n)
        assert np.array_equal(np.isnan(DIJ), np.isnan(want)) and np.allclose(DIJ, want, rtol=1e-14, atol=0, equal_nan=True)
        DI = rng.normal(0, 20, img.shape)
        DI[rng.random(img.shape) < 0.05] = np.nan
        for rt in range(5):
            rho = io.robust_error_function(DI, 7.0, rt)
            np.testing.assert_allclose(rho, orc.robust_error_function(DI, 7.0, rt), rtol=1e-13)
            np.testing.assert_allclose(io.rhop(np.array([0.0, 3.0, 49.0, 1e4]), 7.0, rt),
                                       orc.rhop(np.array([0.0, 3.0, 49.0, 1e4]), 7.0, rt), rtol=1e-14)
        rho = io.robust_error_function(DI, 7.0, 3)
        np.testing.assert_allclose(de.hessian(DIJ), orc.hessian(DIJ), rtol=1e-11)
        np.testing.assert_allclose(de.hessian_robust(DIJ, rho, n), orc.hessian_robust(DIJ, rho), rtol=1e-11)
        np.testing.assert_allclose(io.independent_vector(DIJ, DI, n), orc.independent_vector(DIJ, DI), rtol=1e-10)
        np.testing.assert_allclose(io.independent_vector_robust(DIJ, DI, rho, n), orc.independent_vector_robust(DIJ, DI, rho),
                                   rtol=1e-10)
    cases = [(TransformType.TRANSLATION, [3.25, -1.5]), (TransformType.EUCLIDEAN, [2.0, 1.0, 0.05]),
             (TransformType.SIMILARITY, [1.0, -2.0, 0.02, 0.03]), (TransformType.AFFINITY, [0.5, 0.25, 0.01, 0.02, -0.01, 0.03]),
             (TransformType.HOMOGRAPHY, [0.01, 0.0, 1.0, 0.0, -0.01, 2.0, 1e-5, -2e-5]), (TransformType.AFFINITY, [0.0] * 6)]
    for t, gt in cases:
        got = tr.transform_image(img + 17.0, t, gt)       # min > 0: skimage's "keep cval" clip rule is active
        want = orc.transform_image(img + 17.0, t.value, gt)
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-9)
    with pytest.raises(ValueError):
        io.rhop(np.zeros(3), 1.0, 9)


def test_ipol_warp_matches_reference_golden(nat):
    """``bi.bicubic_interpolation_image`` on the GPU against the reference's own outputs (tests/golden/ipol_warp.npz)."""
    import os
    from inverse_compositional_algorithm_b200 import bicubic_interpolation as bi
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ipol_warp.npz")))
    for i in range(6):
        p, fl = g[f"params_{i}"], g[f"flags_{i}"]
        got = bi.bicubic_interpolation_image(g["image"], p, len(p), bool(fl[0]), int(fl[1]))
        want = g[f"out_{i}"]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-10, equal_nan=True)
    with pytest.raises(ValueError):
        bi.bicubic_interpolation_image(g["image"], np.zeros(5), 5, True, 1)
    assert bi.neumann_bc(-3, 10) == 0 and bi.neumann_bc(12, 10) == 9 and bi.cubic_interpolation([1.0, 2.0, 3.0, 4.0], 0.5) == 2.5


def test_layer_shaped_front_end(nat, rubber_whale):
    """`PyramidalInverseCompositional(...)([I1, I2])`: the batched call shape of the reference's Keras layer, with
    per-pair convergence; every pair must equal its own run of the drop-in driver."""
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import (
        PyramidalInverseCompositional, pyramidal_inverse_compositional_algorithm)
    from inverse_compositional_algorithm_b200.transformation import TransformType
    I2 = rubber_whale["rubber_whale"][100:260, 200:440]
    names = ["rubber_whale_tr", "rubber_whale_eu", "rubber_whale_zo"]
    I1 = np.stack([rubber_whale[n][100:260, 200:440] for n in names])
    layer = PyramidalInverseCompositional(TransformType.SIMILARITY, nscales=3, nu=0.5, TOL=1e-3, robust_type=4,
                                          lambda_=0.0, nanifoutside=True, delta=5)
    p, err, DI, Iw = layer([I1, np.stack([I2] * 3)])
    assert p.shape == (3, 8) and err.shape == (3,) and DI.shape == I1.shape and Iw.dtype == np.float64
    for i in range(3):
        ps, es, DIs, Iws = pyramidal_inverse_compositional_algorithm(I1[i], I2, np.zeros(4), TransformType.SIMILARITY, 3, 0.5,
                                                                     1e-3, 4, 0.0, True, 5, False)
        assert np.array_equal(p[i, :4], ps) and err[i] == es
        assert np.array_equal(np.isnan(Iw[i]), np.isnan(Iws)) and np.nanmax(np.abs(Iw[i] - Iws)) == 0.0
    assert layer.iterations.shape == (3, 3)


def test_device_resident_entry_matches_host_entry(nat):
    """`register_batch_device` (torch CUDA tensors, read in place) gives bit for bit what `register_batch` gives from
    host arrays."""
    import torch
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch, register_batch_device
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType.HOMOGRAPHY
    pairs = [synthetic.make_pair(60 + i, 128, 160, 3, t, max_shift=3.0, margin=32) for i in range(3)]
    I1 = np.stack([a for a, _, _ in pairs]); I2 = np.stack([b for _, b, _ in pairs])
    ph, eh, ih = register_batch(I1, I2, t, nscales=3, robust_type=3, delta=5)
    pd, ed, idv = register_batch_device(torch.from_numpy(I1).cuda(), torch.from_numpy(I2).cuda(), t, nscales=3,
                                        robust_type=3, delta=5)
    assert np.array_equal(pd.cpu().numpy(), ph) and np.array_equal(ed, eh) and np.array_equal(idv, ih)
    with pytest.raises(ValueError):
        register_batch_device(torch.zeros((2, 8, 8, 3)), torch.zeros((2, 8, 8, 3)), t)


def test_zoom_out_matches_oracle(nat, rubber_whale):
    """zm.zoom_out (IPOL-style level; dead code in the reference, parity unpinned): GPU application of the host-built
    operators against the oracle's direct scipy evaluation."""
    from inverse_compositional_algorithm_b200 import zoom as zm
    img = rubber_whale["rubber_whale"][:201, :300].astype(np.float64)
    for f in (0.5, 0.75):
        got = zm.zoom_out(img, f)
        want = orc.zoom_out(img, f)
        assert got.shape == want.shape
        assert rel_err(got, want) <= 1e-5
    gray = zm.zoom_out(img[:, :, :1], 0.5)
    assert rel_err(gray, orc.zoom_out(img[:, :, :1], 0.5)) <= 1e-5
    # strong down-scaling: rows of the operator start 20 samples apart with ~100 taps each -- the border kernels' scalar variant
    full = rubber_whale["rubber_whale"].astype(np.float64)
    assert rel_err(zm.zoom_out(full, 0.05), orc.zoom_out(full, 0.05)) <= 1e-5
    assert rel_err(zm.zoom_out(full, 0.08), orc.zoom_out(full, 0.08)) <= 1e-5


def test_random_shapes_and_options_match_oracle(nat):
    """Seeded sweep over image shapes (odd sizes, smaller than a tile, wider than several tiles), channel counts,
    transform types, error functions, scale counts and frame widths: the motion and the iteration counts of every case
    must match the oracle.  Exercises tile/box edges of the iterate kernel and the border logic of the pyramid."""
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.transformation import TransformType
    rng = np.random.default_rng(424242)
    types = list(TransformType)
    worst = 0.0
    for case in range(24):
        h, w = int(rng.integers(24, 230)), int(rng.integers(24, 300))
        c = int(rng.choice([1, 3]))
        t = types[int(rng.integers(0, 5))]
        rtype = int(rng.integers(0, 5))
        nscales = int(rng.integers(1, 4))
        while min(h, w) * 0.5 ** (nscales - 1) < 12:
            nscales -= 1
        delta = int(rng.integers(0, 6))
        I1, I2, _ = synthetic.make_pair(9000 + case, h, w, c, t, max_shift=1.5, max_lin=0.01, margin=24)
        p, err, iters = register_batch(I1[None], I2[None], t, nscales=nscales, robust_type=rtype, delta=delta)
        a, b = (np.repeat(x, 3, 2) if c == 1 else x for x in (I1, I2))
        trace = []
        po, eo, _, _ = orc.ica_pyramidal(a.astype(np.float64), b.astype(np.float64), np.zeros(t.nparams()), t.value, nscales,
                                         0.5, 1e-3, rtype, 0.0, True, delta, trace=trace)
        epe = _epe(p[0, :t.nparams()], po, t.value, w, h)
        worst = max(worst, epe)
        assert epe <= EPE_TOL, (case, h, w, c, t, rtype, nscales, delta, epe)
        assert int(iters[0].sum()) == len(trace), (case, h, w, c, t, rtype, nscales, delta, iters[0], len(trace))
    print("worst EPE over the sweep:", worst)


def test_plan_reuse_with_moving_images_and_types(nat):
    """One plan, several calls with images at different device addresses (the level-0 TMA tensor maps are re-encoded)
    and changing per-pair transform types (moment degree and loop graph change): every call must equal a fresh run."""
    import torch
    from inverse_compositional_algorithm_b200 import _native, synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.transformation import TransformType
    B, h, w = 3, 120, 152
    sets = []
    for k, t in enumerate([TransformType.HOMOGRAPHY, TransformType.SIMILARITY, TransformType.HOMOGRAPHY]):
        pairs = [synthetic.make_pair(1200 + 10 * k + i, h, w, 3, t, max_shift=2.0, margin=32) for i in range(B)]
        sets.append((t, np.stack([a for a, _, _ in pairs]), np.stack([b for _, b, _ in pairs])))
    plan = _native.Plan(batch=B, height=h, width=w, channels=3, nscales=3, nu=0.5, transform_type=TransformType.HOMOGRAPHY.value,
                        robust_type=3, robust_loop=True, lambda_=0.0, tol=1e-3, max_iter=30, delta=5, nanifoutside=True)
    keep = []          # keep every upload alive so that the addresses really differ
    for rep in range(2):
        for t, I1, I2 in sets:
            a, b = torch.from_numpy(I1).cuda(), torch.from_numpy(I2).cuda()
            keep.append((a, b))
            plan.set_transform_types([t.value] * B)
            p = torch.zeros((B, 8), dtype=torch.float64, device="cuda")
            plan.run_device(a.data_ptr(), b.data_ptr(), p.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            got, err, iters = plan.results()
            want, werr, wit = register_batch(I1, I2, t, nscales=3, robust_type=3, delta=5)
            assert np.array_equal(got, want) and np.array_equal(iters, wit)
    plan.close()


def test_drivers_are_thread_safe(nat):
    """Several Python threads calling the drop-in driver with the same configuration share one cached plan: calls are
    serialised and every thread gets its own (correct, bit-identical) answer."""
    import threading
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import pyramidal_inverse_compositional_algorithm
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType.AFFINITY
    pairs = [synthetic.make_pair(1500 + i, 100, 140, 3, t, max_shift=2.0, margin=32) for i in range(4)]
    want = [pyramidal_inverse_compositional_algorithm(a, b, np.zeros(6), t, 2, 0.5, 1e-3, 3, 0.0, True, 5, False)[0] for a, b, _ in pairs]
    got = [[None] * 3 for _ in pairs]

    def work(i):
        a, b, _ = pairs[i]
        for rep in range(3):
            got[i][rep] = pyramidal_inverse_compositional_algorithm(a, b, np.zeros(6), t, 2, 0.5, 1e-3, 3, 0.0, True, 5, False)[0]

    ths = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [x.start() for x in ths]; [x.join() for x in ths]
    for i in range(4):
        for rep in range(3):
            assert np.array_equal(got[i][rep], want[i])


# ------------------------------------------------------------------ BASELINE configs at full size (committed oracle goldens)
@pytest.fixture(scope="module")
def fullsize_runs(golden_dir):
    import os
    return dict(np.load(os.path.join(golden_dir, "fullsize_runs.npz")))


@pytest.mark.parametrize("name", ["c3_sim", "c3_aff", "c4_a", "c4_b", "c2_a", "c2_diverge"])
def test_full_size_runs_match_oracle_goldens(nat, fullsize_runs, name):
    """BASELINE configs 2, 3 and 4 at FULL size (640x480 gray similarity / affinity quadratic; 1024^2 RGB homography
    Geman-McClure with 20 % occlusion; 1024^2 RGB homography Lorentzian) plus a 1024^2 pair whose motion lies outside
    the capture range (30 iterations at every scale): final parameters, per-scale iteration counts and the
    per-iteration |dp| against oracle runs committed by ``oracle/make_golden_fullsize.py`` (inputs are regenerated
    from their seeds and checksummed)."""
    from oracle import make_golden_fullsize as mg
    from inverse_compositional_algorithm_b200 import _native
    g = fullsize_runs
    seed, H, W, C, t, rt, occ, kw = mg.RUNS[name]
    I1, I2, _ = mg.make_inputs(name)
    assert np.array_equal(g[name + "/checksum"], [I1.sum(dtype=np.float64), I2.sum(dtype=np.float64)]), "inputs differ from the golden run's"
    plan = _native.Plan(batch=1, height=H, width=W, channels=C, nscales=mg.NSCALES, nu=mg.NU, transform_type=t.value,
                        robust_type=rt, robust_loop=rt != 0, lambda_=mg.LAMBDA, tol=mg.TOL, max_iter=30, delta=mg.DELTA,
                        nanifoutside=True, gray_as_rgb=(C == 1), record_trajectory=True)
    p, err, iters, _, _ = plan.run_host(I1[None].astype(np.uint8), I2[None].astype(np.uint8))   # 8-bit through the ABI
    traj = plan.trajectory()[0]
    plan.close()
    want = g[name + "/traj"]
    n = t.nparams()
    assert np.array_equal(iters[0], g[name + "/iters"]), (iters[0], g[name + "/iters"])
    assert len(traj) == len(want)
    if name == "c2_diverge":
        # no convergence: MAX_ITER at every scale on both sides; the (chaotic) path is compared while it is stable
        assert iters[0].min() == 30
        np.testing.assert_allclose(traj[:10, 2], want[:10, 2], rtol=1e-3)
        return
    epe = _epe(p[0, :n], g[name + "/p"], t.value, W, H)
    dev = np.max(np.abs(traj[:, 2] - want[:, 2]) / np.maximum(np.abs(want[:, 2]), 1e-3))
    print(name, "EPE vs oracle", epe, "iterations", iters[0].tolist(), "max |dp| deviation (rel. to max(|dp|, 1e-3))", dev)
    assert epe <= EPE_TOL
    np.testing.assert_allclose(traj[:, 2], want[:, 2], rtol=DP_RTOL, atol=DP_ATOL)
    np.testing.assert_allclose(err[0], g[name + "/err"], rtol=DP_RTOL, atol=DP_ATOL)


def test_hessian_b_8192_wide_gray_vs_rowblocked_oracle(nat, golden_dir):
    """One evaluation of the fused kernel at BASELINE config 5's size (8192 x 8192 gray, homography, Lorentzian)
    against the row-blocked float64 oracle (``oracle/make_golden_8192.py``): this is where fp32 per-lane x-moment
    sums (x^4 ~ 4.5e15) would show.  H entry-wise relative to its diagonal scale, dp = H^-1 b relative."""
    import os
    from inverse_compositional_algorithm_b200 import synthetic
    g = dict(np.load(os.path.join(golden_dir, "hb_8192.npz")))
    name = "8192_lorentzian"
    seed, H, W, rt, lam, delta = g[name + "/cfg"]
    I1, I2 = synthetic.make_large_gray_pair(int(seed), int(H), int(W))
    assert np.array_equal(g[name + "/checksum"], [I1.sum(dtype=np.float64), I2.sum(dtype=np.float64)]), "inputs differ from the golden run's"
    Hg, bg = nat.hessian_b(I1, I2, orc.HOMOGRAPHY, g[name + "/p"], int(rt), float(lam), int(delta), True, gray_as_rgb=True)
    Hw, bw, dpw = g[name + "/H"], g[name + "/b"], g[name + "/dp"]
    scale = np.sqrt(np.outer(np.diag(Hw), np.diag(Hw)))
    h_err = float(np.max(np.abs(Hg - Hw) / scale))
    dpg = np.linalg.solve(Hg, bg)
    dp_err = float(np.max(np.abs(dpg - dpw) / np.abs(dpw)))
    # the error of dp that matters to the loop: the displacement it produces over the image domain
    epe = _epe(dpg, dpw, orc.HOMOGRAPHY, int(W), int(H))
    print("8192^2: max |dH|/sqrt(HiiHjj) =", h_err, " max rel |d dp| =", dp_err, " EPE(dp_gpu, dp_oracle) =", epe, "px")
    assert h_err <= 1e-6       # measured 1.6e-8
    assert dp_err <= HB8192_DP_RTOL
    assert epe <= 1e-5         # measured 2.1e-7 px


# ------------------------------------------------------------------ ingest and input synthesis
@pytest.mark.parametrize("dtype", [np.uint8, np.float32, np.float64])
def test_rgb_to_luminance_ingest(nat, dtype):
    """RGB host images registered on their luminance (converted on the device from the uploaded RGB data) give bit for
    bit what registering the luminance computed on the host gives: Y = 0.2125 R + 0.7154 G + 0.0721 B in float64,
    rounded once to float32."""
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch
    from inverse_compositional_algorithm_b200.transformation import TransformType
    t = TransformType.AFFINITY
    pairs = [synthetic.make_pair(70 + i, 120, 160, 3, t, max_shift=3.0, margin=32) for i in range(2)]
    I1 = np.round(np.stack([a for a, _, _ in pairs])).astype(dtype)
    I2 = np.round(np.stack([b for _, b, _ in pairs])).astype(dtype)

    def luma(x):
        x = x.astype(np.float64)
        return ((0.2125 * x[..., 0] + 0.7154 * x[..., 1]) + 0.0721 * x[..., 2]).astype(np.float32)[..., None]
    pl, el, il = register_batch(I1, I2, t, nscales=3, robust_type=3, delta=5, luminance=True)
    pg, eg, ig = register_batch(luma(I1), luma(I2), t, nscales=3, robust_type=3, delta=5)
    assert np.array_equal(pl, pg) and np.array_equal(il, ig) and np.array_equal(el, eg)
    with pytest.raises(ValueError):
        register_batch(I1[..., :1], I2[..., :1], t, luminance=True)


def test_device_generator_matches_numpy_mirror_and_registers(nat):
    """The device-side pair generator (csrc/ica_generate.cu) against its numpy mirror (same counter-based noise, blur,
    resampling): 8-bit values equal except for a handful of rounding ties; generated pairs register to their ground
    truth; a pair's content depends on (seed, pair index) only."""
    import torch
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.inverse_compositional_algorithm import register_batch_device
    from inverse_compositional_algorithm_b200.transformation import TransformType, end_point_error
    t = TransformType.HOMOGRAPHY
    for (H, W, C, occ) in ((96, 128, 3, 0.2), (120, 160, 1, 0.0)):
        I1, I2, p = synthetic.make_batch_device(4, H, W, C, t, seed=3, pair_offset=2, occlusion=occ, margin=32)
        a1, a2, pg = synthetic.make_pair_hash(3, 3, H, W, C, t, occlusion=occ, margin=32)      # = pair 1 of the batch
        d1 = np.abs(I1[1].cpu().numpy() - a1)
        d2 = np.abs(I2[1].cpu().numpy() - a2)
        assert d1.max() <= 1.0 and d2.max() <= 1.0 and (d1 > 0).mean() < 2e-3 and (d2 > 0).mean() < 2e-3
        assert np.allclose(p[1], np.pad(pg, (0, 8 - len(pg))))
        J1, J2, _ = synthetic.make_batch_device(2, H, W, C, t, seed=3, pair_offset=3, occlusion=occ, margin=32)
        assert torch.equal(J1[0], I1[1]) and torch.equal(J2[0], I2[1])       # pair 3 regenerated alone
    I1, I2, p = synthetic.make_batch_device(6, 256, 320, 3, t, seed=11, max_shift=4.0)
    pr, err, it = register_batch_device(I1, I2, t, nscales=3, robust_type=3, delta=5)
    pr = pr.cpu().numpy()
    for i in range(6):
        assert end_point_error(pr[i], p[i], t, 320, 256)[1] <= 0.05


# ------------------------------------------------------------------ IPOL C++ console logs (second, independent goldens)
@pytest.fixture(scope="module")
def ipol_logs(golden_dir):
    import json, os
    with open(os.path.join(golden_dir, "ipol_cpp_logs.json")) as f:
        return [r for r in json.load(f)["runs"] if r["robust"] == 0]


@pytest.mark.parametrize("idx", range(6))
def test_ipol_options_follow_the_cpp_logs(nat, rubber_whale, ipol_logs, idx):
    """The IPOL-faithful options (``zoom_out`` pyramid, warp domain of ``bicubic_interpolation_image``) as
    modes of the registration plan, against the console logs of the IPOL C++ implementation that the reference stores
    in docs/Algortihm Report.md:38-339 (tests/golden/ipol_cpp_logs.json): six quadratic runs, one and three scales,
    translation / euclidean / similarity -- every printed iteration, identical iteration counts.  (The reference's
    default path agrees with these logs at iteration 0 only; the Charbonnier logs are not followed by the reference
    either: its robust Hessian counts out-of-domain pixels)"""
    from inverse_compositional_algorithm_b200 import _native
    r = ipol_logs[idx]
    tt = {2: orc.TRANSLATION, 3: orc.EUCLIDEAN, 4: orc.SIMILARITY}[r["nparams_code"]]
    I1, I2 = rubber_whale[r["I1"]], rubber_whale[r["I2"]]
    plan = _native.Plan(batch=1, height=388, width=584, channels=3, nscales=r["nscales"], nu=0.5, transform_type=tt,
                        robust_type=0, robust_loop=False, lambda_=0.0, tol=1e-3, max_iter=30, delta=r["delta"],
                        nanifoutside=True, record_trajectory=True, ipol_pyramid=True, ipol_warp=True)
    plan.run_host(I1[None], I2[None])
    traj = plan.trajectory()[0]
    plan.close()
    E = r["entries"]
    n = orc.nparams(tt)
    assert len(traj) == len(E), (len(traj), len(E))
    derr = max(abs(row[2] - e["err"]) for row, e in zip(traj, E))
    dp = max(np.abs(row[4:4 + n] - np.array(e["p"])).max() for row, e in zip(traj, E))
    print("log line", r["line"], r["I1"], "scales", r["nscales"], "iterations", len(E), "max |d|Dp||", derr, "max |dp|", dp)
    assert [int(row[0]) for row in traj] == [e["scale"] for e in E]
    # the logs print 6 decimals; the three-scale runs differ from the C++ pyramid in the resampling kernel of zoom_out
    # (B-spline in the reference's zoom.py, Keys in the C++ code): 3e-5 on the first iterations of the coarsest scale
    tol = 2e-6 if r["nscales"] == 1 else 1.5e-4
    assert derr <= tol and dp <= tol


def test_fused_solve_mode_matches_default_loop(nat, monkeypatch):
    """The opt-in fused solve (ICA_FUSE=1: the CTA that finishes a pair's last chunk solves it inside the iterate kernel,
    one launch per iteration, double-buffered work lists) must give what the default {iterate, solve} loop gives (same
    iteration counts, parameters equal up to the grouping of the fp64 partial sums) -- ragged batch, graph loop and
    host-driven loop."""
    from inverse_compositional_algorithm_b200 import _native, synthetic
    from inverse_compositional_algorithm_b200.transformation import TransformType
    types = [TransformType.HOMOGRAPHY, TransformType.AFFINITY, TransformType.SIMILARITY, TransformType.HOMOGRAPHY]
    pairs = [synthetic.make_pair(1700 + i, 150, 200, 3, t, max_shift=3.0, margin=32) for i, t in enumerate(types)]
    I1 = np.stack([a for a, _, _ in pairs]); I2 = np.stack([b for _, b, _ in pairs])

    def run(host_loop):
        plan = _native.Plan(batch=4, height=150, width=200, channels=3, nscales=3, nu=0.5, transform_type=types[0].value,
                            robust_type=3, robust_loop=True, lambda_=0.0, tol=1e-3, max_iter=30, delta=5, nanifoutside=True,
                            host_loop=host_loop)
        plan.set_transform_types([t.value for t in types])
        out = plan.run_host(I1, I2)
        n = plan.last_launch_count()
        plan.close()
        return out[0], out[1], out[2], n
    ref = run(False)
    monkeypatch.setenv("ICA_FUSE", "1")
    for host_loop in (False, True):
        got = run(host_loop)
        # (the fused solve adds the chunk partials in 11 groups, the solve kernel in 16: last-bit differences of the fp64 sums)
        assert np.array_equal(got[2], ref[2])
        np.testing.assert_allclose(got[0], ref[0], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got[1], ref[1], rtol=1e-6)
    assert got[3] < ref[3]      # one launch per iteration instead of two


@pytest.mark.gpu
def test_march_kernel_matches_tile_kernel(nat, monkeypatch):
    """The opt-in column-march K2 (ICA_K2=march, csrc/ica_march.cu: register-resident 5 x 5 bicubic window, one new window
    row per pixel, zero-padded Keys weights, analytic NaN footprint) must register like the default tile kernel: identical
    iteration counts and parameters within fp32 rounding of the moment sums -- RGB and gray, all moment degrees, a ragged
    batch, a shape that is not a multiple of the 32 x 16 tile, and a rotation beyond the window's slack (global path)."""
    from inverse_compositional_algorithm_b200 import _native, synthetic
    from inverse_compositional_algorithm_b200.transformation import TransformType
    cases = [
        (3, 150, 203, [TransformType.HOMOGRAPHY, TransformType.AFFINITY, TransformType.TRANSLATION, TransformType.HOMOGRAPHY], 3, 3.0),
        (1, 131, 97, [TransformType.SIMILARITY, TransformType.EUCLIDEAN, TransformType.HOMOGRAPHY], 3, 2.0),
        (3, 96, 160, [TransformType.EUCLIDEAN], 0, 2.0),
    ]
    for C, H, W, types, rt, shift in cases:
        pairs = [synthetic.make_pair(2300 + i, H, W, C, t, max_shift=shift, margin=24) for i, t in enumerate(types)]
        if len(types) == 1:      # one strongly rotated pair (7 degrees): most pixels leave the register window
            pairs = [synthetic.make_pair(2400, H, W, C, types[0], margin=24, p_gt=np.array([1.5, -2.0, 0.12]))]
        I1 = np.stack([a for a, _, _ in pairs]); I2 = np.stack([b for _, b, _ in pairs])

        def run():
            plan = _native.Plan(batch=len(types), height=H, width=W, channels=C, nscales=3, nu=0.5,
                                transform_type=types[0].value, robust_type=rt, robust_loop=rt != 0, lambda_=0.0, tol=1e-3,
                                max_iter=30, delta=5, nanifoutside=True)
            plan.set_transform_types([t.value for t in types])
            out = plan.run_host(I1, I2)
            plan.close()
            return out[0], out[1], out[2]
        monkeypatch.delenv("ICA_K2", raising=False)
        ref = run()
        monkeypatch.setenv("ICA_K2", "march")
        got = run()
        monkeypatch.delenv("ICA_K2", raising=False)
        assert np.array_equal(got[2], ref[2]), (C, H, W)
        for b, t in enumerate(types):
            n = t.nparams()
            assert _epe(got[0][b, :n], ref[0][b, :n], t.value, W, H) <= 1e-4, (C, H, W, b)
