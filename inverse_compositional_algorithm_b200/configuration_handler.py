"""INI configuration: same sections, keys and returned dict layout as the reference's
``src/configuration_handler.py:5-67`` (inline ``#`` comments are additionally tolerated, which
makes the reference's own root ``config.ini`` parse)."""
from __future__ import annotations

import configparser

from .image_optimisation import RobustErrorFunctionType
from .transformation import TransformType

_SECTIONS = {
    "InverseCompositionalAlgorithm": "inverse_compositional_algorithm",
    "RobustInverseCompositionalAlgorithm": "robust_inverse_compositional_algorithm",
    "PyramidalInverseCompositionalAlgorithm": "pyramidal_inverse_compositional_algorithm",
}


def create_config_file(filename):
    cfg = configparser.ConfigParser()
    base = {"TOL": "1e-3", "transform_type": "EUCLIDEAN", "verbose": "False"}
    cfg["InverseCompositionalAlgorithm"] = dict(base)
    cfg["RobustInverseCompositionalAlgorithm"] = {
        "TOL": "1e-3", "transform_type": "EUCLIDEAN", "robust_type": "CHARBONNIER", "lambda": "0.0",
        "verbose": "False"}
    cfg["PyramidalInverseCompositionalAlgorithm"] = {
        "TOL": "1e-3", "transform_type": "EUCLIDEAN", "pyramid_levels": "2", "nu": "0.5",
        "robust_type": "QUADRATIC", "lambda": "0.0", "verbose": "False"}
    with open(filename, "w") as fh:
        cfg.write(fh)


def read_config_file(filename):
    cfg = configparser.ConfigParser(inline_comment_prefixes=("#",))
    cfg.read(filename)
    out = {}
    for section, key in _SECTIONS.items():
        sec = cfg[section]
        d = {"TOL": float(sec["TOL"]), "transform_type": TransformType[sec["transform_type"].strip()]}
        if "robust_type" in sec:
            d["robust_type"] = RobustErrorFunctionType[sec["robust_type"].strip()]
            d["lambda"] = float(sec["lambda"])
        if "pyramid_levels" in sec:
            d["pyramid_levels"] = int(sec["pyramid_levels"])
            d["nu"] = float(sec["nu"])
        d["verbose"] = cfg.getboolean(section, "verbose")
        out[key] = d
    return out
