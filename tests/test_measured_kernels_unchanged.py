"""The kernels whose timings and ncu data are committed under profiles/ (round 2: r2m_*, r2n_*, r2l_*) are the kernels of THIS build: every SASS body
listed in profiles/r2_sass_hotpath_unchanged.txt still exists, instruction for instruction, in the objects of hybrid-rendering_b200/build
(tools/sass_function_hashes.py).  A change to a measured kernel makes this fail until it is re-measured and the list regenerated — the
post-budget work of round 2 (TEX instantiations, post-processing, path tracer) was written under this guard."""
import json
import os
import re
import shutil
import subprocess
import sys

import pytest

import pyhr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_measured_kernel_bodies_are_in_the_build(tmp_path):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    objs = [f for f in os.listdir(pyhr.BUILD_DIR) if f.endswith(".o")]
    if not objs:
        pytest.skip("object files not kept next to the library")
    out = tmp_path / "hashes.json"
    env = dict(os.environ, PATH=os.path.dirname(cuobjdump) + os.pathsep + os.environ.get("PATH", ""))
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_function_hashes.py"), pyhr.BUILD_DIR, str(out)], check=True, stdout=subprocess.DEVNULL, env=env)
    have = {(v[0], v[1]) for v in json.load(open(out)).values()}
    listed = []
    for ln in open(os.path.join(ROOT, "profiles", "r2_sass_hotpath_unchanged.txt")):
        m = re.match(r"\s{2}(\S+)\s+(\d+)\s+([0-9a-f]{16})\s*$", ln)
        if m:
            listed.append((m.group(1), int(m.group(2)), m.group(3)))
    assert len(listed) >= 100, "the list of measured kernels could not be read"
    missing = [name for name, n, h in listed if (n, h) not in have]
    assert not missing, f"{len(missing)} measured kernels changed since they were profiled: {missing[:5]}"
