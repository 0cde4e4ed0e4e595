#!/usr/bin/env python
"""Generate the committed golden fixtures under ``tests/golden/``.

TEST INFRASTRUCTURE ONLY; runs in the build container, where ``/root/reference`` exists
(the reference is Python and cannot travel to the GPU box, so its outputs are committed
as fixtures together with this script).

Produces
* ``notebook_trajectories.json`` -- the per-iteration ``|Dp|``/``p``/``lambda`` lines the reference's
  authors stored in ``test/inverse_compositional_algorithm_robust.ipynb`` and
  ``test/inverse_compositional_algorithm.ipynb`` (transcribed verbatim from the cell outputs);
* ``rubber_whale_u8.npz`` -- the five ``test/data/rubber_whale*.png`` inputs of those notebook
  runs as uint8 arrays (data, not source);
* ``reference_runs.npz`` -- trajectories and final parameters of the UNMODIFIED reference
  sources (behind ``oracle/refshim``) on small seeded synthetic pairs, covering every
  transform type and error function, plus helper known-answers (Jacobian, update_transform,
  zoom_in_parameters, warp, rescale).

Usage:  python oracle/make_golden.py [--only notebooks|images|runs]
"""
from __future__ import annotations

import argparse
import contextlib
import io as _io
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_ROOT = os.path.dirname(os.environ.get("ICA_REFERENCE_SRC", "/root/reference/src"))

_LINE = re.compile(
    r"^(?:Iteration (?P<it>\d+): )?\|Dp\|=(?P<err>[^:]+): p=\((?P<p>[^)]*)\)(?:, lambda_=(?P<lam>\S+))?")


def parse_notebook(path):
    """-> list of runs; each run = {cell, sample, first_line, entries:[{scale,it,err,p,lam}]}"""
    nb = json.load(open(path))
    runs = []
    for ci, cell in enumerate(nb["cells"]):
        if cell["cell_type"] != "code":
            continue
        text = "".join("".join(o["text"]) for o in cell.get("outputs", [])
                       if o.get("output_type") == "stream" and o.get("name") == "stdout")
        run, scale = None, -1
        for line in text.splitlines():
            m = re.match(r"^Processing dataset image:\s+(\S+)", line)
            if m:
                run = {"cell": ci, "sample": m.group(1), "entries": []}
                runs.append(run)
                scale = -1
                continue
            m = re.match(r"^Scale: (\d+)", line)
            if m:
                scale = int(m.group(1))
                continue
            m = _LINE.match(line)
            if m and run is not None:
                run["entries"].append({
                    "scale": scale,
                    "it": int(m.group("it")) if m.group("it") else None,
                    "err": float(m.group("err")),
                    "p": [float(v) for v in m.group("p").split()],
                    "lam": float(m.group("lam")) if m.group("lam") else None,
                })
    return [r for r in runs if r["entries"]]


def make_notebooks():
    out = {}
    for name in ("inverse_compositional_algorithm_robust.ipynb",
                 "inverse_compositional_algorithm.ipynb"):
        out[name] = parse_notebook(os.path.join(REF_ROOT, "test", name))
        print(name, [(r["cell"], r["sample"], len(r["entries"])) for r in out[name]])
    with open(os.path.join(GOLDEN, "notebook_trajectories.json"), "w") as f:
        json.dump({"source": "reference test/*.ipynb stored cell outputs (stdout streams)",
                   "notebooks": out}, f, indent=0)


def make_images():
    from PIL import Image
    arrs = {}
    for stem in ("rubber_whale", "rubber_whale_tr", "rubber_whale_rt", "rubber_whale_eu",
                 "rubber_whale_zo"):
        arrs[stem] = np.asarray(Image.open(os.path.join(REF_ROOT, "test", "data", stem + ".png")))
        assert arrs[stem].dtype == np.uint8 and arrs[stem].shape == (388, 584, 3)
    np.savez_compressed(os.path.join(GOLDEN, "rubber_whale_u8.npz"), **arrs)


# ------------------------------------------------------------------ reference runs
RUN_CASES = [
    # name, seed, H, W, transform, robust, nscales, lambda_, occlusion
    ("tr_quad_2s", 1, 72, 96, "TRANSLATION", "QUADRATIC", 2, 0.0, 0.0),
    ("eu_quad_2s", 2, 72, 96, "EUCLIDEAN", "QUADRATIC", 2, 0.0, 0.0),
    ("si_quad_3s", 3, 96, 128, "SIMILARITY", "QUADRATIC", 3, 0.0, 0.0),
    ("af_quad_3s", 4, 96, 128, "AFFINITY", "QUADRATIC", 3, 0.0, 0.0),
    ("ho_quad_3s", 5, 96, 128, "HOMOGRAPHY", "QUADRATIC", 3, 0.0, 0.0),
    ("tr_char_2s", 6, 72, 96, "TRANSLATION", "CHARBONNIER", 2, 0.0, 0.0),
    ("eu_lore_3s", 7, 96, 128, "EUCLIDEAN", "LORENTZIAN", 3, 0.0, 0.0),
    ("si_gmcc_3s", 8, 96, 128, "SIMILARITY", "GERMAN_MCCLURE", 3, 0.0, 0.0),
    ("af_char_3s", 9, 97, 131, "AFFINITY", "CHARBONNIER", 3, 0.0, 0.0),
    ("ho_lore_3s", 10, 96, 128, "HOMOGRAPHY", "LORENTZIAN", 3, 0.0, 0.0),
    ("ho_gmcc_occ_3s", 11, 128, 128, "HOMOGRAPHY", "GERMAN_MCCLURE", 3, 0.0, 0.2),
    ("af_trunc_2s", 12, 72, 96, "AFFINITY", "TRUNCATED_QUADRATIC", 2, 0.0, 0.0),
    ("ho_char_lam_2s", 13, 80, 112, "HOMOGRAPHY", "CHARBONNIER", 2, 12.5, 0.0),
    ("ho_lore_odd_4s", 14, 203, 251, "HOMOGRAPHY", "LORENTZIAN", 4, 0.0, 0.0),
]


def _patched_rhop(io_mod):
    """The reference's TRUNCATED_QUADRATIC branch applies ``if`` to an array and raises
    (image_optimisation.py:40).  Patch that ONE branch to its element-wise reading so the case
    can be pinned; every other branch is the reference's own code."""
    orig = io_mod.rhop

    def rhop(t2, lambda_, type_):
        if type_ == io_mod.RobustErrorFunctionType.TRUNCATED_QUADRATIC:
            return np.where(t2 < lambda_ * lambda_, np.ones_like(t2), np.zeros_like(t2))
        return orig(t2, lambda_, type_)
    return rhop


def _run_verbose(fn, **kw):
    buf = _io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = fn(**kw)
    entries, scale = [], -1
    for line in buf.getvalue().splitlines():
        m = re.match(r"^Scale: (\d+)", line)
        if m:
            scale = int(m.group(1))
            continue
        m = _LINE.match(line)
        if m:
            entries.append((scale, float(m.group("err")), [float(v) for v in m.group("p").split()],
                            float(m.group("lam")) if m.group("lam") else np.nan))
    return res, entries


def make_runs():
    from oracle import reference_loader
    from inverse_compositional_algorithm_b200 import synthetic
    from inverse_compositional_algorithm_b200.transformation import TransformType as MyT
    ref = reference_loader.load()
    ica, tr, io, de, zm, bi = (ref[k] for k in ("ica", "tr", "io", "de", "zm", "bi"))
    io.rhop = _patched_rhop(io)
    out = {}
    names = []
    for (name, seed, H, W, tname, rname, nscales, lam, occ) in RUN_CASES:
        ttype = tr.TransformType[tname]
        rtype = io.RobustErrorFunctionType[rname]
        max_shift = 0.35 * 2 ** (nscales + 1)
        I1, I2, p_gt = synthetic.make_pair(seed, H, W, 3, MyT[tname], max_shift=max_shift,
                                           occlusion=occ, margin=32)
        (p, err, DI, Iw), entries = _run_verbose(
            ica.pyramidal_inverse_compositional_algorithm,
            I1=I1.astype(np.float64), I2=I2.astype(np.float64), p=np.zeros(ttype.nparams()),
            transform_type=ttype, nscales=nscales, nu=0.5, TOL=1e-3, robust_type=rtype,
            lambda_=lam, nanifoutside=True, delta=5, verbose=True)
        n = ttype.nparams()
        traj = np.full((len(entries), 3 + 8), np.nan)
        for i, (s, e, pp, l) in enumerate(entries):
            traj[i, 0], traj[i, 1], traj[i, 2] = s, e, l
            traj[i, 3:3 + n] = pp
        names.append(name)
        out[name + "/cfg"] = np.array([seed, H, W, ttype.value, rtype.value, nscales, lam, occ,
                                       max_shift, 5], dtype=np.float64)
        out[name + "/p"] = np.asarray(p, dtype=np.float64)
        out[name + "/p_gt"] = p_gt
        out[name + "/err"] = np.float64(err)
        out[name + "/traj"] = traj
        out[name + "/input_sum"] = np.array([I1.astype(np.float64).sum(), I2.astype(np.float64).sum()])
        # sparse probes of the returned images (positions fixed; NaN pattern included)
        yy = np.linspace(0, H - 1, 9).astype(int)
        xx = np.linspace(0, W - 1, 11).astype(int)
        out[name + "/Iw_probe"] = Iw[np.ix_(yy, xx)]
        out[name + "/DI_probe"] = DI[np.ix_(yy, xx)]
        out[name + "/nan_count"] = np.int64(np.isnan(Iw).sum())
        print(name, "iters", len(entries), "err", err, "epe-ish |p-p_gt|", np.abs(p - p_gt).max())
    out["names"] = np.array(names)

    # ---- helper known-answers from the reference
    rng = np.random.default_rng(7)
    for t in tr.TransformType:
        n = t.nparams()
        out[f"jac/{t.name}"] = de.jacobian(t, 5, 4)
        ps, dps, res, zres = [], [], [], []
        for _ in range(6):
            p = rng.uniform(-0.05, 0.05, n)
            dp = rng.uniform(-0.05, 0.05, n)
            if t.name == "HOMOGRAPHY":
                p[[2, 5]] *= 100; dp[[2, 5]] *= 100; p[[6, 7]] *= 1e-2; dp[[6, 7]] *= 1e-2
            else:
                p[:2] *= 100; dp[:2] *= 100
            ps.append(p.copy()); dps.append(dp.copy())
            res.append(tr.update_transform(p.copy(), dp.copy(), t))
            zres.append(zm.zoom_in_parameters(p.copy(), t, 97.0, 49.0, 194.0, 97.0))
        out[f"upd/{t.name}/p"] = np.array(ps)
        out[f"upd/{t.name}/dp"] = np.array(dps)
        out[f"upd/{t.name}/out"] = np.array(res)
        out[f"zoomin/{t.name}/out"] = np.array(zres)
        out[f"p2m/{t.name}"] = tr.params2matrix(ps[0], t)
    # warp + rescale on a small seeded image (through the reference's own call sites)
    img = (rng.uniform(0, 255, (37, 53, 3))).astype(np.float32).astype(np.float64)
    out["warp/img"] = img
    for t, p in (("TRANSLATION", [1.3, -2.6]), ("EUCLIDEAN", [0.7, 1.9, 0.03]),
                 ("SIMILARITY", [-1.2, 0.4, 0.02, -0.015]),
                 ("AFFINITY", [2.2, -1.1, 0.01, 0.02, -0.015, 0.012]),
                 ("HOMOGRAPHY", [0.01, 0.02, 2.2, -0.015, 0.012, -1.1, 2e-4, -1e-4]),
                 ("HOMOGRAPHY", [0.0] * 8)):
        key = f"warp/{t}/{'id' if not any(p) else 'p'}"
        out[key + "/p"] = np.array(p)
        out[key + "/out"] = bi.bicubic_interpolation_skimage(img, np.array(p), tr.TransformType[t],
                                                            True, 5)
    from skimage.transform import rescale  # the shim, exactly as ica.py:333-336 calls it
    for shape in ((37, 53, 3), (64, 48, 3), (97, 131, 3)):
        im = rng.uniform(0, 255, shape).astype(np.float32).astype(np.float64)
        out[f"rescale/{shape[0]}x{shape[1]}/in"] = im
        out[f"rescale/{shape[0]}x{shape[1]}/out"] = rescale(
            im, 0.5, mode='constant', cval=0, order=3, anti_aliasing=True, channel_axis=2,
            preserve_range=True)
    np.savez_compressed(os.path.join(GOLDEN, "reference_runs.npz"), **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", choices=["notebooks", "images", "runs"])
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    if a.only in (None, "notebooks"):
        make_notebooks()
    if a.only in (None, "images"):
        make_images()
    if a.only in (None, "runs"):
        make_runs()
