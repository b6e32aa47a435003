"""The reference-side seam of INTEGRATION.md section 1, exercised against the reference checkout when it is present (the
build container; the GPU box has no /root/reference): the reference's own `Model` / `parse_model` builds with the head
class names rebound to the B200 drop-ins, the modules carry the builder's tags and reference checkpoints load strictly."""
import os
import subprocess
import sys

import pytest

REF = os.environ.get("YC_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "nets")), reason="reference checkout not present")
def test_reference_model_builds_with_rebound_heads():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_integration_worker.py"), REF], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "INTEGRATION_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
