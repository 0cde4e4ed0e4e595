#!/usr/bin/env python
"""Transcribes the console logs of the IPOL C++ implementation that the reference stores in
``docs/Algortihm Report.md`` (quadratic runs :38-339, Charbonnier runs :348-433) into
``tests/golden/ipol_cpp_logs.json``: per run the command line options and every printed
``Iteration k: |Dp|=...: p=(...)`` / ``|Dp|=...: p=(...), lambda=...`` line with its scale.

    python oracle/make_golden_ipol_logs.py

These are a SECOND, independent set of golden trajectories (another implementation, in another language, by the
algorithm's authors): the reference's default path only agrees with them at iteration 0 (it uses skimage's warp domain),
the IPOL options of SURVEY 8f-4 (``warp_mode="ipol"``, ``pyramid_mode="ipol"``) follow them to the printed digits.
"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/docs/Algortihm Report.md"

CMD = re.compile(r"^\./inverse_compositional_algorithm \./data2/(\w+)\.png \./data2/(\w+)\.png -t (\d+) -r (\d+) .*-d (\d+) .*-n ?(\d*)")
SCALE = re.compile(r"^Scale: (\d+)")
ITQ = re.compile(r"^Iteration (\d+): \|Dp\|=([-\d.eE+]+): p=\(([^)]*)\)")
ITR = re.compile(r"^\|Dp\|=([-\d.eE+]+): p=\(([^)]*)\)(?:, lambda=([-\d.eE+]+))?")


def main():
    runs, cur, scale = [], None, 0
    with open(SRC) as f:
        for ln, line in enumerate(f, 1):
            line = line.strip()
            m = CMD.match(line)
            if m:
                cur = {"line": ln, "I1": m.group(1), "I2": m.group(2), "nparams_code": int(m.group(3)),
                       "robust": int(m.group(4)), "delta": int(m.group(5)), "nscales": int(m.group(6) or 1), "entries": []}
                runs.append(cur)
                scale = 0
                continue
            if line.startswith("###") or line.startswith("## "):
                cur = None
                continue
            if cur is None:
                continue
            m = SCALE.match(line)
            if m:
                scale = int(m.group(1))
                continue
            m = ITQ.match(line)
            if m:
                cur["entries"].append({"scale": scale, "it": int(m.group(1)), "err": float(m.group(2)),
                                       "p": [float(v) for v in m.group(3).split()]})
                continue
            m = ITR.match(line)
            if m:
                cur["entries"].append({"scale": scale, "it": None, "err": float(m.group(1)),
                                       "p": [float(v) for v in m.group(2).split()],
                                       "lam": float(m.group(3)) if m.group(3) else None})
    runs = [r for r in runs if r["entries"]]
    out = {"source": "docs/Algortihm Report.md of mfournigault/inverse_compositional_algorithm (IPOL C++ console logs)",
           "note": "-t is the number of parameters (2 translation, 3 euclidean, 4 similarity); -r 0 quadratic, 4 Charbonnier",
           "runs": runs}
    path = os.path.join(ROOT, "tests", "golden", "ipol_cpp_logs.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    for r in runs:
        print(r["line"], r["I1"], "t", r["nparams_code"], "r", r["robust"], "scales", r["nscales"], len(r["entries"]), "lines")


if __name__ == "__main__":
    main()
