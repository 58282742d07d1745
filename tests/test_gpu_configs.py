"""GPU parity at the sizes BASELINE.json's configs name, every file against the UNMODIFIED reference
binary (oracle/_ref), time-boxed:
  configs[0]  C1  4.6 Mbp genome, 30x 100 bp SE single-line FASTA, K=31, -p 8
  configs[1]  C2 shape at 10 Mbp (3.3e4 x the toy cases; the full 100 Mbp cmp is scripts/cli_full.sh, about 6 minutes of reference time)
  configs[4]  C5  20 Mbp genome, three libraries (FASTQ PE, FASTA PE, reverse_seq SE FASTA+FASTQ), K=63 -p 8 -a 4 -R, then the
              reference's own `contig -R` on both outputs and a cmp of .contig/.Arc/.updated.edge/.ContigIndex
Set PGB200_SKIP_CONFIG_TESTS=1 to skip them (each needs 1-4 minutes of reference CPU time)."""
import os
import subprocess

import pytest

from soapdenovo2_b200 import api, synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need():
    if os.environ.get("PGB200_SKIP_CONFIG_TESTS"):
        pytest.skip("PGB200_SKIP_CONFIG_TESTS set")
    if not util.have_ref():
        pytest.skip("oracle/_ref not shipped")


def _engine(cfg, out, K, P, extra, env=None):
    r = subprocess.run([api.BIN63, "pregraph", "-s", cfg, "-K", str(K), "-p", str(P), "-o", out, *extra], capture_output=True, text=True,
                       env=dict(os.environ, **(env or {})), timeout=1200)
    assert r.returncode == 0, r.stderr[-4000:]
    return r.stderr


def test_c1_ecoli_sized_se_fasta_k31(tmp_path):
    cfg = synth.config_c1(str(tmp_path))
    ref, gpu = str(tmp_path / "ref"), str(tmp_path / "gpu")
    util.run([util.REF63, "pregraph", "-s", cfg, "-K", "31", "-p", "8", "-a", "1", "-o", ref], timeout=1200)
    _engine(cfg, gpu, 31, 8, ("-a", "1"))
    util.compare(ref, gpu, util.SUFFIXES)


def test_c2_shape_10mbp_pe_fastq_k63(tmp_path):
    cfg = synth.config_c2(str(tmp_path))
    ref, gpu = str(tmp_path / "ref"), str(tmp_path / "gpu")
    util.run([util.REF63, "pregraph", "-s", cfg, "-K", "63", "-p", "8", "-a", "2", "-R", "-o", ref], timeout=1500)
    _engine(cfg, gpu, 63, 8, ("-a", "2", "-R"), env={"PGB200_CHUNK_MB": "64", "PGB200_SKM": "1"})   # aggregated pass 1, several chunks
    util.compare(ref, gpu, util.SUFFIXES_R)


def test_c5_three_libraries_20mbp_with_contig_diff(tmp_path):
    cfg = synth.config_c5(str(tmp_path))
    ref, gpu = str(tmp_path / "ref"), str(tmp_path / "gpu")
    util.run([util.REF63, "pregraph", "-s", cfg, "-K", "63", "-p", "8", "-a", "4", "-R", "-o", ref], timeout=1800)
    _engine(cfg, gpu, 63, 8, ("-a", "4", "-R"))
    util.compare(ref, gpu, util.SUFFIXES_R)
    for pre in (ref, gpu):
        util.run([util.REF63, "contig", "-g", pre, "-R"], timeout=1800)
    util.compare(ref, gpu, ["contig", "Arc", "updated.edge", "ContigIndex"])
