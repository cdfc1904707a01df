"""Fixture generator (run in the build container only): the reference's OWN ``script/data_preprocess.py``
``slate2trajectory`` and ``data_augment`` on synthetic page records -> tests/golden/ingest_pages.json.

    python tests/golden/make_ingest_golden.py
"""
import json
import runpy
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402
from rl4rs_b200 import synth  # noqa: E402


def main():
    ref_harness.install_stubs()
    script = os.path.join(ref_harness.REFERENCE_ROOT, "script", "data_preprocess.py")

    def run_stage(stage, src, dst):          # the script is its own CLI: `data_preprocess.py <in> <out> <stage>` (:169-180)
        argv = sys.argv
        sys.argv = [script, src, dst, stage]
        try:
            runpy.run_path(script, run_name="__main__")
        finally:
            sys.argv = argv
    cat = synth.make_catalog()
    log = synth.make_log(28, pages=1, catalog=cat, keep_hist=True, seed=77)
    recs = synth.render_records(log, cat)
    # 7 sessions x 4 pages (same session id / user fields, sequence_id 1..4); session 5 gets only 2 pages for data_augment
    pages = []
    for s in range(7):
        base = recs[4 * s].split("@")
        for p in range(4):
            f = recs[4 * s + p].split("@")
            pages.append("@".join([str(1000 * s + p), base[1], str(p + 1), f[3], f[4], base[5], base[6], f[7], base[8]]))
    header = "timestamp@session_id@sequence_id@exposed_items@user_feedback@user_seqfeature@user_protrait@item_feature@behavior_policy_id"
    full = [header] + pages + [""]
    short = [header] + [p for i, p in enumerate(pages) if i not in (18, 19, 11)] + [""]    # sessions with 2 and 3 pages
    out = {"pages": full, "short": short}
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, "in.csv"), os.path.join(d, "out.csv")
        open(src, "w").write("\n".join(full))
        run_stage("slate2trajectory", src, dst)
        out["trajectories"] = open(dst).read().split("\n")
        open(src, "w").write("\n".join(short))
        np.random.seed(5)
        run_stage("data_augment", src, dst)
        out["augmented"] = open(dst).read().split("\n")
    with open(os.path.join(ROOT, "tests", "golden", "ingest_pages.json"), "w") as f:
        json.dump(out, f)
    print("trajectories:", len([x for x in out["trajectories"] if x]), "augmented lines:", len([x for x in out["augmented"] if x]))


if __name__ == "__main__":
    main()
