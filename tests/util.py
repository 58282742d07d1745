"""Shared helpers for the parity tests: run the reference binary / the C model / the engine and compare files."""
import filecmp
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF63 = os.path.join(ROOT, "oracle", "_ref", "SOAPdenovo-63mer")
REF127 = os.path.join(ROOT, "oracle", "_ref", "SOAPdenovo-127mer")
MODEL63 = os.path.join(ROOT, "oracle", "pregraph_model_63")
MODEL127 = os.path.join(ROOT, "oracle", "pregraph_model_127")
ENGINE63 = os.path.join(ROOT, "soapdenovo2_b200", "bin", "pregraph-b200-63mer")
ENGINE127 = os.path.join(ROOT, "soapdenovo2_b200", "bin", "pregraph-b200-127mer")
SUFFIXES = ["kmerFreq", "vertex", "preGraphBasic", "preArc", "edge.gz"]
SUFFIXES_R = SUFFIXES + ["markOnEdge", "path"]


def have_ref():
    return os.path.exists(REF63) and os.path.exists(REF127)


def build_oracle():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "model"], check=True)


def run(cmd, timeout=600):
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired as ex:   # show where it hung
        err = ex.stderr.decode(errors="replace") if isinstance(ex.stderr, bytes) else (ex.stderr or "")
        raise AssertionError(f"{' '.join(cmd)}\ntimed out after {timeout} s; stderr tail:\n{err[-3000:]}")
    assert r.returncode == 0, f"{' '.join(cmd)}\nrc={r.returncode}\n{r.stderr[-4000:]}"
    return r.stderr


def run_ref(binary, cfg, out, K, P, extra=()):
    return run([binary, "pregraph", "-s", cfg, "-K", str(K), "-p", str(P), "-o", out, *extra])


def run_model(binary, cfg, out, K, P, extra=()):
    return run([binary, "-s", cfg, "-K", str(K), "-p", str(P), "-o", out, *extra])


def compare(a, b, suffixes):
    bad = [s for s in suffixes if not filecmp.cmp(f"{a}.{s}", f"{b}.{s}", shallow=False)]
    assert not bad, f"files differ: {bad} ({a} vs {b})"
