"""The literal drop-in: the reference's callers bind by ``sys.path.append("../src/")`` followed by top-level
imports (``/root/reference/test/inverse_compositional_algorithm_robust.ipynb:49-51``,
``/root/reference/test/test_derivatives.py:7-9``).  ``src_dropin/`` is a directory of flat modules with the
reference's module names; these tests run the reference's own import lines and asserts against it, each in a
fresh interpreter so that the flat module names never leak into the rest of the suite.
"""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "src_dropin")
REF_TEST = "/root/reference/test/test_derivatives.py"


def _run(code, cwd, timeout=600):
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=cwd, env=env, capture_output=True,
                       text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout + "\n" + r.stderr
    return r.stdout


@pytest.fixture()
def caller_dir(tmp_path):
    """``<tmp>/test`` next to ``<tmp>/src`` -> src_dropin: the layout the reference's callers assume."""
    os.symlink(DROPIN, tmp_path / "src")
    d = tmp_path / "test"
    d.mkdir()
    return str(d)


def test_reference_import_lines_work_unchanged(caller_dir):
    out = _run("""
        import os, sys
        sys.path.append(os.path.abspath("../src/"))
        from inverse_compositional_algorithm import inverse_compositional_algorithm, robust_inverse_compositional_algorithm, pyramidal_inverse_compositional_algorithm
        import configuration_handler as cfh
        import image_optimisation as io
        import transformation as tf
        from derivatives import hessian, jacobian, TransformType
        import bicubic_interpolation as bi, zoom as zm, constants as cts, derivatives as de, transformation as tr
        assert TransformType is tf.TransformType and tf.TransformType.HOMOGRAPHY.nparams() == 8
        assert io.RobustErrorFunctionType.CHARBONNIER.value == 4 and cts.MAX_ITER == 30
        for mod, names in ((tr, "update_transform project params2matrix matrix2params transform_image"),
                           (io, "rhop robust_error_function independent_vector independent_vector_robust parametric_solve steepest_descent_images"),
                           (de, "jacobian hessian hessian_robust inverse_hessian"),
                           (bi, "bicubic_interpolation_skimage bicubic_interpolation_image neumann_bc cubic_interpolation"),
                           (zm, "zoom_size zoom_in_parameters zoom_out"), (cfh, "create_config_file read_config_file")):
            for n in names.split():
                assert callable(getattr(mod, n)), (mod.__name__, n)
        cfh.create_config_file("config.ini")
        params = cfh.read_config_file("config.ini")
        assert set(params) == {"inverse_compositional_algorithm", "robust_inverse_compositional_algorithm",
                               "pyramidal_inverse_compositional_algorithm"}
        print("ok", sorted(params["pyramidal_inverse_compositional_algorithm"]))
        """, caller_dir)
    assert out.startswith("ok")


def test_reference_jacobian_known_answers_through_dropin(caller_dir):
    """The asserts of the reference's ``TestJacobian`` (test_derivatives.py:13-68), restated, through the unchanged
    import line of that file."""
    _run("""
        import os, sys
        import numpy as np
        sys.path.append(os.path.abspath("../src/"))
        from derivatives import hessian, jacobian, TransformType
        want = {
            TransformType.TRANSLATION: [[[1, 0, 0, 1], [1, 0, 0, 1]], [[1, 0, 0, 1], [1, 0, 0, 1]]],
            TransformType.EUCLIDEAN: [[[1, 0, 0, 0, 1, 0], [1, 0, 0, 0, 1, 1]], [[1, 0, -1, 0, 1, 0], [1, 0, -1, 0, 1, 1]]],
            TransformType.SIMILARITY: [[[1, 0, 0, 0, 0, 1, 0, 0], [1, 0, 1, 0, 0, 1, 0, 1]],
                                       [[1, 0, 0, -1, 0, 1, 1, 0], [1, 0, 1, -1, 0, 1, 1, 1]]],
            TransformType.AFFINITY: [[[1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0], [1, 0, 1, 0, 0, 0, 0, 1, 0, 0, 1, 0]],
                                     [[1, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1], [1, 0, 1, 1, 0, 0, 0, 1, 0, 0, 1, 1]]],
        }
        for t, w in want.items():
            J = jacobian(t, 2, 2)
            assert J.shape == (2, 2, 2 * t.nparams())
            np.testing.assert_array_almost_equal(J, np.array(w, dtype=float))
        """, caller_dir)


@pytest.mark.skipif(not os.path.exists(REF_TEST), reason="the reference tree is only present in the build container")
def test_reference_unittest_file_runs_against_dropin(caller_dir):
    """The reference's own, unmodified ``test/test_derivatives.py::TestJacobian`` executed from a ``test/``
    directory whose sibling ``src/`` is this repository's ``src_dropin/``."""
    r = subprocess.run([sys.executable, REF_TEST, "TestJacobian"], cwd=caller_dir, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stderr


@pytest.mark.gpu
def test_reference_hessian_tests_through_dropin(caller_dir):
    """``TestHessianFunction.test_valid_dij`` / ``test_empty_dij`` of the reference (test_derivatives.py:73-82, 97-101):
    H == sum of outer products, on the GPU helper behind the flat module."""
    _run("""
        import os, sys
        import numpy as np
        sys.path.append(os.path.abspath("../src/"))
        from derivatives import hessian, jacobian, TransformType
        rng = np.random.default_rng(0)
        DIJ = rng.random((3, 3, 4, 2))
        want = np.zeros((2, 2))
        for i in range(3):
            for j in range(3):
                for c in range(4):
                    want += np.outer(DIJ[i, j, c], DIJ[i, j, c])
        np.testing.assert_array_almost_equal(hessian(DIJ), want)
        np.testing.assert_array_almost_equal(hessian(np.zeros((0, 0, 3, 2))), np.zeros((2, 2)))
        """, caller_dir)


@pytest.mark.gpu
def test_notebook_cell15_replayed_with_reference_imports(caller_dir, notebook_runs):
    """robust.ipynb cells 2, 5, 7-10, 15 with their own import lines and keyword call, on the images the notebook
    reads from disk (committed as arrays); final parameters against the lines the notebook stored."""
    golden = os.path.join(ROOT, "tests", "golden")
    out = _run(f"""
        import os, sys, json
        import numpy as np
        sys.path.append(os.path.abspath("../src/"))
        from inverse_compositional_algorithm import inverse_compositional_algorithm, robust_inverse_compositional_algorithm, pyramidal_inverse_compositional_algorithm
        import configuration_handler as cfh
        import image_optimisation as io
        import transformation as tf
        with open("config.ini", "w") as f:       # the reference's test/config.ini
            f.write('''[InverseCompositionalAlgorithm]
        tol = 1e-3
        transform_type = EUCLIDEAN
        verbose = False

        [RobustInverseCompositionalAlgorithm]
        tol = 1e-3
        transform_type = EUCLIDEAN
        robust_type = CHARBONNIER
        lambda = 0.0
        verbose = False

        [PyramidalInverseCompositionalAlgorithm]
        tol = 1e-3
        transform_type = EUCLIDEAN
        pyramid_levels = 3
        nu = 0.5
        robust_type = CHARBONNIER
        lambda = 0.0
        verbose = True
        '''.replace("        ", ""))
        params = cfh.read_config_file("config.ini")
        params_pica = params["pyramidal_inverse_compositional_algorithm"]
        imgs = np.load(os.path.join({golden!r}, "rubber_whale_u8.npz"))
        dataset_tu = {{
            "rubber_whale_tr": (tf.TransformType.TRANSLATION, "rubber_whale_tr"),
            "rubber_whale_rt": (tf.TransformType.EUCLIDEAN, "rubber_whale_rt"),
            "rubber_whale_eu": (tf.TransformType.EUCLIDEAN, "rubber_whale_eu"),
            "rubber_whale_zo": (tf.TransformType.SIMILARITY, "rubber_whale_zo"),
        }}
        res = {{}}
        for sample_key, (transformation_type, fname) in dataset_tu.items():
            original_image = imgs[fname]                 # the notebook swaps the pair (cell 10)
            transformed_image = imgs["rubber_whale"]
            p = np.zeros(transformation_type.nparams())
            p, error, DI, Iw = pyramidal_inverse_compositional_algorithm(
                I1=original_image, I2=transformed_image, p=p, transform_type=transformation_type,
                nscales=params_pica["pyramid_levels"], nu=params_pica["nu"], TOL=params_pica["TOL"],
                robust_type=params_pica["robust_type"], lambda_=params_pica["lambda"], nanifoutside=True, delta=10,
                verbose=params_pica["verbose"])
            assert DI.shape == original_image.shape and Iw.dtype == np.float64
            res[sample_key] = dict(p=[float(v) for v in p], error=float(error))
        print("RESULT" + json.dumps(res))
        """, caller_dir)
    import json
    import numpy as np
    from oracle import ica_oracle as orc
    res = json.loads(out.split("RESULT")[-1])
    ttypes = {"rubber_whale_tr": orc.TRANSLATION, "rubber_whale_rt": orc.EUCLIDEAN, "rubber_whale_eu": orc.EUCLIDEAN,
              "rubber_whale_zo": orc.SIMILARITY}
    for sample, tt in ttypes.items():
        entries = [r for r in notebook_runs["inverse_compositional_algorithm_robust.ipynb"]
                   if r["cell"] == 15 and r["sample"] == sample][0]["entries"]
        assert orc.end_point_error(np.array(res[sample]["p"]), np.array(entries[-1]["p"]), tt, 584, 388)[1] <= 1e-3
        np.testing.assert_allclose(res[sample]["error"], entries[-1]["err"], rtol=2e-3)
    # verbose=True printed the reference's per-scale / per-iteration lines
    assert "Scale: 2" in out and "|Dp|=" in out and "lambda_=" in out
