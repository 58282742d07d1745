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
