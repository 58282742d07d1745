"""CPU: known-answer vectors (tests/golden/kats.json, minted from the reference sources -- SURVEY.md section 4) against the
engine's own host+device helpers (kmer.cuh, engine_impl.cuh) compiled into a host harness, and the zlib CRC identity."""
import json
import os
import struct
import subprocess
import zlib

from tests import util

GOLD = json.load(open(os.path.join(util.ROOT, "tests", "golden", "kats.json")))


def _harness(tmp_path):
    exe = str(tmp_path / "host_kat")
    subprocess.run(["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                    os.path.join(util.ROOT, "tests", "host_kat.cu")], check=True, capture_output=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    return {l.split()[0]: l.split()[1:] for l in out.splitlines()}


def test_engine_helpers_match_reference_kats(tmp_path):
    got = _harness(tmp_path)
    assert got["hash_zero_63"] == [GOLD["hash_zero_63"]]
    assert got["kmer63_fwd"] == GOLD["kmer63_fwd"] and got["kmer63_rc"] == GOLD["kmer63_rc"]
    assert int(got["kmer63_smaller_fwd_rc"][0]) == GOLD["kmer63_smaller_fwd_rc"]
    assert got["kmer63_hash_fwd"] == [GOLD["kmer63_hash_fwd"]] and got["kmer63_hash_rc"] == [GOLD["kmer63_hash_rc"]]
    assert got["kmer63_hash_fwd_mer127"] == [GOLD["kmer63_hash_fwd"]]        # identical in the 127-mer build
    assert got["rolling_rc_matches"] == ["1"]
    for i, (req, lf, size, mx) in enumerate(GOLD["init_kmerset"]):
        assert [int(x) for x in got[f"init_kmerset_{i}"]] == [size, mx]
    assert got["static_set_size_a1_p3_63"] == ["16777259"]                      # `-p 3 -a 1` static size
    assert got["sizeof_slot"] == ["32", "64"]
    assert got["rc128_quirk"][0] == "1"


def test_crc_identity_with_zlib():
    """hash = sign_extend32(zlib.crc32(pack('<QQ', high, low), 0xFFFFFFFF))."""
    hi, lo = (int(x, 16) for x in GOLD["kmer63_fwd"])
    assert zlib.crc32(struct.pack("<QQ", hi, lo), 0xFFFFFFFF) == int(GOLD["kmer63_hash_fwd"], 16)
    assert zlib.crc32(struct.pack("<QQ", 0, 0), 0xFFFFFFFF) == 0xFFFFFFFF


def test_base_codes():
    for ch, code in GOLD["base2int"].items():
        assert (ord(ch) & 6) >> 1 == code and (ord(ch.lower()) & 6) >> 1 == code


def test_oracle_model_matches_reference_kats():
    """The CPU restatement's own primitives against the same golden vectors (both builds)."""
    util.build_oracle()
    for exe in (util.MODEL63, util.MODEL127):
        out = subprocess.run([exe, "-V"], check=True, capture_output=True, text=True).stdout
        got = {l.split()[0]: l.split()[1:] for l in out.splitlines()}
        assert got["hash_zero"] == [GOLD["hash_zero_63"]]
        assert got["kmer63_fwd"] == GOLD["kmer63_fwd"] and got["kmer63_rc"] == GOLD["kmer63_rc"]
        assert int(got["kmer63_smaller_fwd_rc"][0]) == GOLD["kmer63_smaller_fwd_rc"]
        assert got["kmer63_hash_fwd"] == [GOLD["kmer63_hash_fwd"]] and got["kmer63_hash_rc"] == [GOLD["kmer63_hash_rc"]]
        for i, (req, lf, size, mx) in enumerate(GOLD["init_kmerset"]):
            assert [int(x) for x in got[f"init_kmerset_{i}"]] == [size, mx]
