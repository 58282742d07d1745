"""Link-level drop-in: the reference's own main() / contig / map / scaff objects linked
against libpregraph_b200.so (scripts/link_dropin.sh: the six replaced files left out, pregraph_shim.c in their place, flavour
fixed at link time).  `pregraph`, `contig` and the whole `all` pipeline through that binary must write the same files as the
unmodified reference binary."""
import os
import subprocess

import pytest

from soapdenovo2_b200 import synth
from tests import util

pytestmark = pytest.mark.gpu
B63 = os.path.join(util.ROOT, "oracle", "_ref", "SOAPdenovo-63mer-b200")
B127 = os.path.join(util.ROOT, "oracle", "_ref", "SOAPdenovo-127mer-b200")


@pytest.fixture(scope="module", autouse=True)
def _need():
    if not (util.have_ref() and os.path.exists(B63) and os.path.exists(B127)):
        pytest.skip("oracle/_ref (reference objects + linked drop-in binaries) not shipped")


def test_pregraph_and_contig_through_the_reference_main(tmp_path):
    cfg = synth.scenario_multilib(str(tmp_path))
    ref, gpu = str(tmp_path / "ref"), str(tmp_path / "gpu")
    for exe, out in ((util.REF63, ref), (B63, gpu)):
        util.run([exe, "pregraph", "-s", cfg, "-K", "63", "-p", "8", "-a", "1", "-R", "-o", out])
        util.run([exe, "contig", "-g", out, "-R"])
    util.compare(ref, gpu, util.SUFFIXES_R + ["contig", "Arc", "updated.edge", "ContigIndex"])


def test_127mer_flavour_is_fixed_at_link_time(tmp_path):
    cfg = synth.scenario_pe_fastq(str(tmp_path))
    ref, gpu = str(tmp_path / "ref"), str(tmp_path / "gpu")
    env = dict(os.environ)
    for exe, out in ((util.REF127, ref), (B127, gpu)):
        r = subprocess.run([exe, "pregraph", "-s", cfg, "-K", "91", "-p", "4", "-a", "1", "-o", out], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
    util.compare(ref, gpu, util.SUFFIXES)


def test_all_pipeline_in_one_process(tmp_path):
    """`all` = pregraph + contig + map + scaff in ONE process (main.c:117-545): the stages after the GPU pregraph must find
    everything they need (they re-read K from .preGraphBasic and take -p from their own argv)."""
    cfg = synth.scenario_pe_fastq(str(tmp_path), genome_len=30000, n_pairs=4000)
    outs = {}
    for tag, exe in (("ref", util.REF63), ("gpu", B63)):
        out = str(tmp_path / tag)
        util.run([exe, "all", "-s", cfg, "-K", "31", "-p", "1", "-a", "1", "-R", "-o", out], timeout=240)
        outs[tag] = out
    util.compare(outs["ref"], outs["gpu"], util.SUFFIXES_R + ["contig", "Arc", "updated.edge", "ContigIndex", "scafSeq", "scaf", "links", "newContigIndex"])


def test_stage_writes_the_edge_sidecar_and_contig_reads_it(tmp_path):
    """f2: with PGB200_EDGE_SIDECAR=1 the GPU stage also writes <prefix>.edge.b200; it must equal what the host converter makes from
    the stage's own (byte-identical) .edge.gz, and the reference's `contig`, linked with csrc/contig_sidecar.c, must build the same
    contigs from it WITHOUT the .edge.gz."""
    import shutil
    from soapdenovo2_b200 import api
    cfg = synth.scenario_multilib(str(tmp_path))
    ref, gpu = str(tmp_path / "ref"), str(tmp_path / "gpu")
    util.run([util.REF63, "pregraph", "-s", cfg, "-K", "63", "-p", "8", "-a", "1", "-R", "-o", ref])
    r = subprocess.run([B63, "pregraph", "-s", cfg, "-K", "63", "-p", "8", "-a", "1", "-R", "-o", gpu], capture_output=True, text=True,
                       env=dict(os.environ, PGB200_EDGE_SIDECAR="1"), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    util.compare(ref, gpu, util.SUFFIXES_R)
    side = open(gpu + ".edge.b200", "rb").read()
    shutil.copy(gpu + ".edge.b200", gpu + ".edge.b200.stage")
    api.edge_gz_to_sidecar(gpu, 63, 0)
    assert open(gpu + ".edge.b200", "rb").read() == side
    os.remove(gpu + ".edge.gz")
    util.run([util.REF63, "contig", "-g", ref, "-R"])
    util.run([B63, "contig", "-g", gpu, "-R"])
    util.compare(ref, gpu, ["contig", "Arc", "updated.edge", "ContigIndex"])
    # "only": the sidecar without the (sequential, slow) deflate of the edge text; everything else unchanged
    only = str(tmp_path / "only")
    r = subprocess.run([B63, "pregraph", "-s", cfg, "-K", "63", "-p", "8", "-a", "1", "-R", "-o", only], capture_output=True, text=True,
                       env=dict(os.environ, PGB200_EDGE_SIDECAR="only"), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert not os.path.exists(only + ".edge.gz")
    assert open(only + ".edge.b200", "rb").read() == side
    util.compare(ref, only, [x for x in util.SUFFIXES_R if x != "edge.gz"])
