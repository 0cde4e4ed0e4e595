"""Pin the CPU oracle (oracle/ica_oracle.py) against the reference's own stored outputs.

(a) per-iteration trajectories printed in the reference's notebooks
    (tests/golden/notebook_trajectories.json, transcribed by oracle/make_golden.py from
    test/inverse_compositional_algorithm_robust.ipynb and test/inverse_compositional_algorithm.ipynb);
(b) the Jacobian known-answers of the reference's test/test_derivatives.py:13-68;
(c) outputs of the unmodified reference sources run behind oracle/refshim
    (tests/golden/reference_runs.npz).
CPU only; no GPU, no /root/reference needed.
"""
import numpy as np
import pytest

from oracle import ica_oracle as orc
from inverse_compositional_algorithm_b200 import synthetic
from inverse_compositional_algorithm_b200.transformation import TransformType

NB_ROBUST = "inverse_compositional_algorithm_robust.ipynb"
NB_QUAD = "inverse_compositional_algorithm.ipynb"
RTOL = 1e-9  # oracle vs stored lines at <= 1e-9 relative

SAMPLE = {  # notebook dataset_tu: sample -> (file suffix, transform, gt used by transform_image)
    "rubber_whale_tr": ("tr", orc.TRANSLATION, [10, 5]),
    "rubber_whale_rt": ("rt", orc.EUCLIDEAN, [0.0, 0.0, -0.1]),
    "rubber_whale_eu": ("eu", orc.EUCLIDEAN, [10.0, 5.0, -0.1]),
    "rubber_whale_zo": ("zo", orc.SIMILARITY, [0.0, 0.0, -0.1, 0.0]),
}


def _runs(notebook_runs, nb, cell):
    return {r["sample"]: r["entries"] for r in notebook_runs[nb] if r["cell"] == cell}


def _compare(trace, entries, with_scale):
    assert len(trace) == len(entries), (len(trace), len(entries))
    for got, want in zip(trace, entries):
        if with_scale:
            s, _, err, p, lam = got
            assert s == want["scale"]
        else:
            _, err, p, lam = got
        np.testing.assert_allclose(err, want["err"], rtol=RTOL)
        np.testing.assert_allclose(p, want["p"], rtol=RTOL, atol=1e-12)
        if want["lam"] is not None:
            np.testing.assert_allclose(lam, want["lam"], rtol=1e-13)


def test_robust_single_scale_translation_charbonnier(rubber_whale, notebook_runs):
    """robust.ipynb cell 13 (stored lines :305-335): 18 iterations, lambda schedule."""
    entries = _runs(notebook_runs, NB_ROBUST, 13)["rubber_whale_tr"]
    trace = []
    p, err, DI, Iw = orc.ica_robust(rubber_whale["rubber_whale_tr"], rubber_whale["rubber_whale"],
                                    np.zeros(2), orc.TRANSLATION, 1e-3, orc.CHARBONNIER, 0.0,
                                    True, 10, trace=trace)
    _compare(trace, entries, with_scale=False)
    assert len(trace) == 18
    assert np.isnan(Iw[0, 0, 0]) and np.isfinite(Iw[200, 300, 0])


@pytest.mark.parametrize("sample", list(SAMPLE))
def test_pyramidal_charbonnier_from_disk(sample, rubber_whale, notebook_runs):
    """robust.ipynb cell 15 (stored lines :566-785): 3 scales, inputs read from disk."""
    entries = _runs(notebook_runs, NB_ROBUST, 15)[sample]
    suffix, ttype, _ = SAMPLE[sample]
    trace = []
    orc.ica_pyramidal(rubber_whale["rubber_whale_" + suffix], rubber_whale["rubber_whale"],
                      np.zeros(orc.nparams(ttype)), ttype, 3, 0.5, 1e-3, orc.CHARBONNIER, 0.0,
                      True, 10, trace=trace)
    _compare(trace, entries, with_scale=True)


def test_quadratic_single_scale_translation_in_memory(rubber_whale, notebook_runs):
    """ipynb cell 14 (stored lines :319-348): I1 = transform_image(rubber_whale, [10,5])."""
    entries = _runs(notebook_runs, NB_QUAD, 14)["rubber_whale_tr"]
    I2 = rubber_whale["rubber_whale"]
    I1 = orc.transform_image(I2, orc.TRANSLATION, [10, 5])
    trace = []
    orc.ica_quadratic(I1, I2, np.zeros(2), orc.TRANSLATION, 1e-3, True, 10, trace=trace)
    _compare(trace, entries, with_scale=False)
    assert len(trace) == 13


@pytest.mark.parametrize("sample", ["rubber_whale_eu", "rubber_whale_zo"])
def test_pyramidal_quadratic_in_memory(sample, rubber_whale, notebook_runs):
    """ipynb cell 16 (stored lines :566-748): 3 scales, QUADRATIC, in-memory bilinear inputs."""
    entries = _runs(notebook_runs, NB_QUAD, 16)[sample]
    _, ttype, gt = SAMPLE[sample]
    I2 = rubber_whale["rubber_whale"]
    I1 = orc.transform_image(I2, ttype, gt)
    trace = []
    orc.ica_pyramidal(I1, I2, np.zeros(orc.nparams(ttype)), ttype, 3, 0.5, 1e-3, orc.QUADRATIC, 0,
                      True, 10, trace=trace)
    _compare(trace, entries, with_scale=True)


# ------------------------------------------------------------------ (b) Jacobian known-answers
def test_jacobian_known_answers():
    """Values of the reference's test/test_derivatives.py:13-68 on a 2x2 grid."""
    J = orc.jacobian(orc.TRANSLATION, 2, 2)
    assert J.shape == (2, 2, 4)
    np.testing.assert_array_equal(J[1, 1], [1, 0, 0, 1])
    J = orc.jacobian(orc.EUCLIDEAN, 2, 2)
    np.testing.assert_array_equal(J[1, 1], [1, 0, -1, 0, 1, 1])
    np.testing.assert_array_equal(J[0, 1], [1, 0, 0, 0, 1, 1])
    J = orc.jacobian(orc.SIMILARITY, 2, 2)
    np.testing.assert_array_equal(J[1, 0], [1, 0, 0, -1, 0, 1, 1, 0])
    J = orc.jacobian(orc.AFFINITY, 2, 2)
    np.testing.assert_array_equal(J[1, 1], [1, 0, 1, 1, 0, 0, 0, 1, 0, 0, 1, 1])


def test_hessian_matches_pixel_loop():
    """test/test_derivatives.py:73-82 (test_valid_dij): H = sum_ij DIJ_ij^T DIJ_ij."""
    rng = np.random.default_rng(0)
    DIJ = rng.random((3, 3, 4, 2))
    want = sum(DIJ[i, j].T @ DIJ[i, j] for i in range(3) for j in range(3))
    np.testing.assert_allclose(orc.hessian(DIJ), want, rtol=1e-13)
    assert orc.hessian(np.zeros((0, 0, 3, 2))).shape == (2, 2)


# ------------------------------------------------------- (c) unmodified reference, committed runs
def _names(reference_runs):
    return [str(n) for n in reference_runs["names"]]


def test_reference_helper_known_answers(reference_runs):
    g = reference_runs
    for t in TransformType:
        np.testing.assert_array_equal(orc.jacobian(t.value, 5, 4), g[f"jac/{t.name}"])
        for p, dp, want, zwant in zip(g[f"upd/{t.name}/p"], g[f"upd/{t.name}/dp"],
                                      g[f"upd/{t.name}/out"], g[f"zoomin/{t.name}/out"]):
            np.testing.assert_allclose(orc.update_transform(p.copy(), dp, t.value), want,
                                       rtol=1e-13, atol=1e-15)
            np.testing.assert_allclose(
                orc.zoom_in_parameters(p, t.value, 97.0, 49.0, 194.0, 97.0), zwant, rtol=1e-15)
        np.testing.assert_allclose(orc.params2matrix(g[f"upd/{t.name}/p"][0], t.value),
                                   g[f"p2m/{t.name}"], rtol=1e-15)
    for key in [k for k in g if k.startswith("warp/") and k.endswith("/out")]:
        tname = key.split("/")[1]
        got = orc.warp_bicubic(g["warp/img"], g[key[:-4] + "/p"], TransformType[tname].value)
        np.testing.assert_allclose(got, g[key], rtol=1e-13, equal_nan=True)
    for key in [k for k in g if k.startswith("rescale/") and k.endswith("/out")]:
        got = orc.build_pyramid(g[key[:-4] + "/in"], 2, 0.5)[1]
        np.testing.assert_allclose(got, g[key], rtol=1e-13)


@pytest.mark.parametrize("idx", range(14))
def test_reference_runs_match_oracle(idx, reference_runs):
    g = reference_runs
    name = _names(g)[idx]
    seed, H, W, tt, rt, nscales, lam, occ, max_shift, delta = g[name + "/cfg"]
    I1, I2, p_gt = synthetic.make_pair(int(seed), int(H), int(W), 3, TransformType(int(tt)),
                                       max_shift=float(max_shift), occlusion=float(occ), margin=32)
    np.testing.assert_allclose([I1.astype(np.float64).sum(), I2.astype(np.float64).sum()],
                               g[name + "/input_sum"], rtol=1e-12)  # generator did not drift
    trace = []
    p, err, DI, Iw = orc.ica_pyramidal(I1, I2, np.zeros(orc.nparams(int(tt))), int(tt),
                                       int(nscales), 0.5, 1e-3, int(rt), float(lam), True,
                                       int(delta), trace=trace)
    traj = g[name + "/traj"]
    assert len(trace) == len(traj)
    n = orc.nparams(int(tt))
    for (s, _, e, pp, l), row in zip(trace, traj):
        assert s == row[0]
        np.testing.assert_allclose(e, row[1], rtol=1e-7)
        np.testing.assert_allclose(pp, row[3:3 + n], rtol=1e-7, atol=1e-11)
    np.testing.assert_allclose(p, g[name + "/p"], rtol=1e-7, atol=1e-11)
    assert int(np.isnan(Iw).sum()) == int(g[name + "/nan_count"])


def test_ipol_warp_matches_reference_golden():
    """``bicubic_interpolation_image`` (the IPOL-style warp, secondary API): the oracle's restatement against outputs
    of the unmodified reference (numba, fastmath) stored by ``oracle/make_golden_ipol.py``."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ipol_warp.npz")))
    for i in range(6):
        p, fl = g[f"params_{i}"], g[f"flags_{i}"]
        got = orc.bicubic_interpolation_image(g["image"], p, len(p), bool(fl[0]), int(fl[1]))
        want = g[f"out_{i}"]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-10, equal_nan=True)


def test_oracle_ipol_modes_follow_the_cpp_logs(rubber_whale):
    """Pins the oracle's IPOL options (``warp_mode="ipol"``, ``pyramid_mode="ipol"``: bicubic_interpolation_image's domain,
    zoom.zoom_out levels) with an INDEPENDENT implementation: the IPOL C++ console logs the reference stores in
    docs/Algortihm Report.md:38-339 (transcribed by oracle/make_golden_ipol_logs.py).  Single scale: every printed digit;
    three scales: same iteration counts, 3e-5 (the C++ zoom interpolates with Keys, zoom.py with a B-spline)."""
    import json, os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ipol_cpp_logs.json")) as f:
        runs = [r for r in json.load(f)["runs"] if r["robust"] == 0]
    assert len(runs) == 6
    for r in runs[:3]:          # (the other three take a minute of CPU; the GPU suite runs all six)
        tt = {2: orc.TRANSLATION, 3: orc.EUCLIDEAN, 4: orc.SIMILARITY}[r["nparams_code"]]
        I1 = rubber_whale[r["I1"]].astype(np.float64)
        I2 = rubber_whale[r["I2"]].astype(np.float64)
        trace = []
        orc.ica_pyramidal(I1, I2, np.zeros(orc.nparams(tt)), tt, r["nscales"], 0.5, 1e-3, orc.QUADRATIC, 0.0, True, r["delta"],
                          trace=trace, warp_mode="ipol", pyramid_mode="ipol")
        E = r["entries"]
        assert len(trace) == len(E)
        tol = 1e-6 if r["nscales"] == 1 else 1e-4
        for t, e in zip(trace, E):
            assert t[0] == e["scale"] and abs(t[2] - e["err"]) <= tol and np.abs(np.array(t[3]) - np.array(e["p"])).max() <= tol
