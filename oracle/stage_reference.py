#!/usr/bin/env python
"""Stage the UNMODIFIED reference package next to the oracle (test / measurement infrastructure only).

    python oracle/stage_reference.py            # copies /root/reference/code/src -> oracle/_ref/src

The reference is pure Python, so "building" it is a file copy of the modules on the hot path:
`src/{__init__,losses,trainer}.py`, `src/models/*.py`, `src/utils/{__init__,trainer_utils}.py`.  `oracle/_ref/` is
git-ignored (the reference's sources never enter this repo's history) but not gpurun-ignored, so the staged copy travels
to the GPU box, where `bench.py --impl reference` and the `eager_gpu_baseline` leg import it as the same-box comparator
(`oracle/ref_runner.py`).  Nothing under `clear_vae_b200/` may import it.  A `MANIFEST.json` with the sha256 of every
staged file is written so a reader can check that the copy is byte-identical to the reference.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["__init__.py", "losses.py", "trainer.py", "models/__init__.py", "models/vae.py", "models/mi_estimator.py", "models/cnn.py",
         "utils/__init__.py", "utils/trainer_utils.py"]


def stage(reference_root="/root/reference", verbose=False) -> bool:
    src = os.path.join(reference_root, "code", "src")
    if not os.path.isdir(src):
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))   # GPU box: use what travelled with the snapshot
    manifest = {}
    for rel in FILES:
        a, b = os.path.join(src, rel), os.path.join(DST, "src", rel)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        if not os.path.exists(a):
            if rel.endswith("__init__.py"):
                open(b, "w").close()
                continue
            raise FileNotFoundError(a)
        shutil.copyfile(a, b)
        manifest[rel] = hashlib.sha256(open(b, "rb").read()).hexdigest()
    json.dump(dict(source=src, files=manifest), open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"staged {len(manifest)} reference files into {DST}")
    return True


if __name__ == "__main__":
    ok = stage(*(sys.argv[1:2]), verbose=True)
    sys.exit(0 if ok else 1)
