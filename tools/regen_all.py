#!/usr/bin/env python
"""Re-emit every generated solution listed in yask_b200/csrc/gen/manifest.json (needs the stencil compiler under tools/_refc,
i.e. the build container) and rewrite the registry include once at the end."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
man = json.load(open(os.path.join(ROOT, "yask_b200", "csrc", "gen", "manifest.json")))
names = sorted(man)
for i, name in enumerate(names):
    m = man[name]
    cmd = [sys.executable, "-m", "yask_b200.emitter.yask_cuda_emit", "--stencil", m["stencil"], "--elem-bytes",
           str(m["elem_bytes"]), "--name", name]
    if m.get("radius"):
        cmd += ["--radius", str(m["radius"])]
    if i + 1 < len(names):
        cmd.append("--no-registry")
    subprocess.run(cmd, check=True, cwd=ROOT, stdout=subprocess.DEVNULL)
print(f"re-emitted {len(names)} solutions")
