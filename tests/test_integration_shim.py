"""The reference-side binding (integration/libp_b200_shim.hpp) must compile against the reference's own headers.
Needs the scratch copy of the reference made by oracle/refbuild/build_ref.sh (OCCA's generated headers), so it
runs in the build container only and is skipped elsewhere."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORK = os.environ.get("LIBP_REF_WORK", "/tmp/libp_ref")


@pytest.mark.skipif(not os.path.isdir(os.path.join(WORK, "occa", "include")) or not os.path.isdir("/usr/local/cuda/include"),
                    reason="reference scratch tree (oracle/refbuild/build_ref.sh) not present")
def test_shim_compiles_against_reference_headers():
    p = subprocess.run(["bash", os.path.join(ROOT, "integration", "check_shim.sh")], capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    assert "shim compiles" in p.stdout
