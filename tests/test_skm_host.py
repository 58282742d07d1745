"""CPU: the super-k-mer partition logic of the aggregated pass 1 (soapdenovo2_b200/csrc/skm.cuh) compiled for the host and checked
against a naive per-position restatement of the instance rules: runs tile every read, bucket assignment is
strand-symmetric, the self-contained run records reproduce every (k-mer, left, right, rank) instance, the lane-packing map, payload_merge of partial aggregates equals applying
all instances in read order (including saturation).  K = 13..127, both table widths.  See tests/host_skm.cu."""
import os
import subprocess

from tests import util


def test_super_kmer_partition_and_merge_on_host(tmp_path):
    exe = str(tmp_path / "host_skm")
    subprocess.run(["nvcc", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                    os.path.join(util.ROOT, "tests", "host_skm.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "ALL OK"
    assert len(lines) == 11 and all("errors=0" in l for l in lines[:-1])
    assert any("saturated=" in l and "saturated=0" not in l for l in lines[:-1])   # the saturation case really saturates
