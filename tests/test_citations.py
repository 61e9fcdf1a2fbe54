"""Every `hpc/...:line` / `python/...:line` citation in the sources, headers, tests and design documents must point into an
existing file of the reference tree and inside its length.  Runs only where /root/reference exists (the build container);
the GPU box has no reference tree."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PAT = re.compile(r"\b((?:hpc|python)/[\w/\.]+\.(?:c|h|py|md)):(\d+)(?:-(\d+))?")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_reference_citations_point_at_existing_lines():
    files = [os.path.join(ROOT, f) for f in ("DESIGN.md", "INTEGRATION.md", "README.md", "bench.py", "__graft_entry__.py")]
    for sub in ("include", "image-processing-graph-laplacian_b200", "oracle", "tests"):
        files += glob.glob(os.path.join(ROOT, sub, "**", "*"), recursive=True)
    lengths, bad, checked = {}, [], 0
    for f in files:
        if not os.path.isfile(f) or not f.endswith((".h", ".c", ".cu", ".cuh", ".py", ".md")) or os.sep + "build" + os.sep in f:
            continue
        text = open(f, errors="replace").read()
        for m in PAT.finditer(text):
            path, a, b = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            full = os.path.join(REF, path)
            checked += 1
            if not os.path.exists(full):
                bad.append((os.path.relpath(f, ROOT), m.group(0), "no such file in the reference"))
                continue
            if path not in lengths:
                lengths[path] = sum(1 for _ in open(full, errors="replace"))
            if a < 1 or b < a or b > lengths[path]:
                bad.append((os.path.relpath(f, ROOT), m.group(0), f"the file has {lengths[path]} lines"))
    assert checked > 200
    assert not bad, bad
