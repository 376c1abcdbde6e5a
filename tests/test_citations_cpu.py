"""Every `file:line` citation of the reference in the boundary header, the docs, the oracle and the engine sources must
point at an existing reference file and a valid line range (build container only: needs /root/reference)."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PAT = re.compile(r"((?:/root/reference/)?(?:src|include|utils|docs)/[\w./-]+\.(?:cpp|hpp|h|mk|pl|py|md)|[\w-]+\.(?:cpp|hpp)):(\d+)(?:-(\d+))?")
OWN = {os.path.basename(p) for p in glob.glob(os.path.join(ROOT, "yask_b200", "csrc", "*")) + glob.glob(os.path.join(ROOT, "oracle", "*"))}


def test_reference_citations_resolve():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present")
    index = {}
    for root, _, fs in os.walk(REF):
        for f in fs:
            index.setdefault(f, []).append(os.path.join(root, f))
    files = ["include/yask_b200.h", "DESIGN.md", "INTEGRATION.md", "oracle/yask_oracle.c", "oracle/oracle.py", "oracle/ref_driver.cpp",
             "yask_b200/include/yask_kernel_api.hpp"]
    for pat in ("yask_b200/csrc/*.cu", "yask_b200/csrc/*.cuh", "yask_b200/csrc/*.cpp", "yask_b200/csrc/*.h", "yask_b200/*.py",
                "yask_b200/emitter/*.py"):
        files += [os.path.relpath(p, ROOT) for p in glob.glob(os.path.join(ROOT, pat))]
    bad, n = [], 0
    for fn in files:
        for m in PAT.finditer(open(os.path.join(ROOT, fn)).read()):
            path, l0, l1 = m.group(1), int(m.group(2)), int(m.group(3) or m.group(2))
            if "/" not in path and path in OWN:
                continue          # a citation of one of this repo's own files
            n += 1
            if path.startswith(REF):
                cands = [path]
            elif "/" in path:
                cands = [os.path.join(REF, path)]
            else:
                cands = index.get(path, [])
            cands = [c for c in cands if os.path.exists(c)]
            if not cands:
                bad.append((fn, m.group(0), "no such reference file"))
                continue
            nl = max(sum(1 for _ in open(c, errors="ignore")) for c in cands)
            if l1 > nl or l0 > l1:
                bad.append((fn, m.group(0), f"file has {nl} lines"))
    assert n > 100 and not bad, bad
