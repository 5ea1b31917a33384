#!/usr/bin/env python3
"""fetch_ref.py -- place the reference's own two files of the hot path where the GPU box can run them.

TEST INFRASTRUCTURE ONLY.  The reference is pure Python (no build step): its "compiled form" is its
source.  This recipe copies ``viterbi_markov.py`` and ``Pd_plotter.py`` byte for byte from
``/root/reference`` (present only in the build container) into ``oracle/_ref/`` -- git-ignored, so the
reference's sources never enter this repository's history, but NOT gpurun-ignored, so the directory
travels to the GPU box like a built ``.so`` does.  ``__graft_entry__.build()`` runs it when the
reference is present; ``oracle/ref_harness.py`` (bench.py's ``--impl reference`` arm and its
``cpu_baseline`` leg) then times the reference's unmodified ``run_experiment`` from there.

    python oracle/fetch_ref.py [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("viterbi_markov.py", "Pd_plotter.py")


def fetch(ref: str = "/root/reference") -> dict:
    """Copy FILES from ``ref`` into oracle/_ref/ and write a manifest of their sha256; no-op (returns {})
    when ``ref`` does not exist (the GPU box)."""
    if not os.path.isdir(ref):
        return {}
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for name in FILES:
        src = os.path.join(ref, name)
        shutil.copyfile(src, os.path.join(DEST, name))
        with open(src, "rb") as f:
            manifest[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "sha256": manifest}, f, indent=1)
    return manifest


def available() -> bool:
    return all(os.path.exists(os.path.join(DEST, name)) for name in FILES)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    got = fetch(ap.parse_args().ref)
    print("fetched" if got else "reference not present; nothing fetched", got)
