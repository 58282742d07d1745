"""CPU-only: the C-ABI library loads and exports every symbol include/pregraph_b200.h declares; host-side logic."""
import os
import re
import subprocess

import pytest

from soapdenovo2_b200 import api
from tests import util


@pytest.fixture(scope="module", autouse=True)
def _build():
    subprocess.run(["make", "-s", "-j8", "-C", os.path.join(util.ROOT, "soapdenovo2_b200", "csrc")], check=True)


def test_exports_match_header():
    hdr = open(os.path.join(util.ROOT, "include", "pregraph_b200.h")).read()
    declared = set(re.findall(r"\b(pgb200_[a-z0-9_]+|call_pregraph)\s*\(", hdr))
    declared -= {"pgb200_engine", "pgb200_params", "pgb200_pass1_stats", "pgb200_graph_stats"}
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)
    lib = api.load()
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_oracle_in_product():
    """The product path must never import / link / exec anything under oracle/."""
    for root, _, files in os.walk(os.path.join(util.ROOT, "soapdenovo2_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".c", ".h")) or f == "Makefile":
                txt = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle/" not in txt and "pregraph_model" not in txt, f"{f} references the oracle"


def test_cli_usage_exits_nonzero():
    r = subprocess.run([api.BIN63, "pregraph"], capture_output=True, text=True)
    assert r.returncode != 0 and "pregraph -s configFile -o outputGraph" in r.stderr


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.EngineError):
        api.PregraphEngine(K=31)


def test_dropin_links_against_reference_objects():
    """The reference's own objects (minus the six replaced files) + pregraph_shim.o + libpregraph_b200.so link into a complete
    SOAPdenovo binary whose call_pregraph is the shim (flavour fixed at link time) and whose engine entry comes from the library."""
    if not os.path.isdir(os.path.join(util.ROOT, "oracle", "_ref", "o63")):
        pytest.skip("oracle/_ref objects absent (built where /root/reference exists)")
    subprocess.run(["bash", os.path.join(util.ROOT, "scripts", "link_dropin.sh")], check=True, capture_output=True)
    for fl in ("63", "127"):
        exe = os.path.join(util.ROOT, "oracle", "_ref", f"SOAPdenovo-{fl}mer-b200")
        syms = subprocess.run(["nm", "-D", "--defined-only", exe], capture_output=True, text=True).stdout
        und = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
        assert "pgb200_pregraph_main" in und
        full = subprocess.run(["nm", exe], capture_output=True, text=True).stdout
        assert " T call_pregraph" in full and " T call_heavygraph" in full and "prlRead2HashTable" not in full
