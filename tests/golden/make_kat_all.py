#!/usr/bin/env python
"""Parses EVERY record of EVERY results file the reference ships (QEC_LDPC/results/**, 300 files) into
tests/golden/kat_all.json: seed, weight, frame count, iteration limit, error probability (from the file name) and the
published counters.  Run in the build container (needs /root/reference); the JSON is committed, the results files are
not.  tests/test_gpu_parity.py::test_all_published_results_files replays the records on the GPU."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
RES = "/root/reference/QEC_LDPC/results"
FIELDS = {"Rand Seed": "seed", "Errors Tested": "count", "Errors With X": "xTested", "Errors With Z": "zTested",
          "Error Weight": "W", "Corrected": "corrected", "Syndrome Errors X": "synX", "Syndrome Errors Z": "synZ",
          "Logical Errors": "logical", "Convergence Fail X": "cvX", "Convergence Fail Z": "cvZ",
          "Duration(micro-s)": "duration_us"}


def main():
    out = []
    for root, _, files in sorted(os.walk(RES)):
        for name in sorted(files):
            if not name.endswith(".txt"):
                continue
            path = os.path.join(root, name)
            m_it = re.search(r"_MAX_(\d+)", name)
            m_p = re.search(r"_p_([0-9.]+?)(?:\.txt|_)", name)
            m_n = re.search(r"n=(\d+)", name)
            recs = []
            for line in open(path, errors="replace").read().splitlines():
                if ":" not in line:
                    continue
                k, v = line.split(":", 1)
                k = k.strip()
                if k == "Code":
                    recs.append({"code_name": v.strip()})
                elif k in FIELDS and recs:
                    try:
                        recs[-1][FIELDS[k]] = int(v)
                    except ValueError:
                        pass
            sub = os.path.relpath(root, RES)
            # `archive/` and the dated directories hold output of earlier versions of the program (other code-name
            # format, no logical-error detection, a different decoder): not reproducible by design.  The files under
            # [4,5,10,61,9,49]/ labelled p_0.01 were produced with errorProbability 0.02.
            current = sub in (".", "[2,3,6,7,2,3]", "[4,5,10,61,9,49]")
            for i, r in enumerate(recs):
                r["group"] = "current" if current else "earlier program version"
                if current and m_p:
                    r["errorProbability"] = 0.02 if sub == "[4,5,10,61,9,49]" else float(m_p.group(1))
                r.update(source=os.path.relpath(path, "/root/reference"), record_index=i,
                         maxit=int(m_it.group(1)) if m_it else None, p_in_name=float(m_p.group(1)) if m_p else None,
                         n=int(m_n.group(1)) if m_n else None)
                out.append(r)
    json.dump(out, open(os.path.join(HERE, "kat_all.json"), "w"), indent=0, sort_keys=True)
    print(len(out), "records from", len({r["source"] for r in out}), "files")


if __name__ == "__main__":
    main()
