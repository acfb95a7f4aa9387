"""The reference's OWN acceptance files, run UNMODIFIED on top of the package under test:

  * test/test_decoder.py (unittest; 10 cases over construction, the single-node debug API, check functions and
    Hamming(7,4) decoding) -- its only missing dependency, the third-party `galois`, is supplied by the stand-in
    in tests/stubs/galois (GF2 arrays with XOR addition);
  * sims/sim_reconciliation.py, the driver script, by path, in its three modes, on a small (3,6) code.

The files are byte-identical copies made by oracle/build_ref.py into oracle/_ref/reference_tests/ (git-ignored,
travels to the GPU box; /root/reference itself does not exist there).  A sha256 of each copy is printed so a
reader can check them against the reference checkout.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "qam-reconciliation_b200")
STAGE = os.path.join(ROOT, "oracle", "_ref", "reference_tests")
STUBS = os.path.join(ROOT, "tests", "stubs")

pytestmark = pytest.mark.gpu


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([PKG, STUBS, env.get("PYTHONPATH", "")])
    return env


def _staged(rel):
    from oracle import build_ref
    build_ref.stage_reference_tests()
    p = os.path.join(STAGE, rel)
    if not os.path.exists(p):
        pytest.skip(f"{rel}: the reference checkout is not reachable and no staged copy exists")
    print(rel, "sha256", hashlib.sha256(open(p, "rb").read()).hexdigest())
    return p


def test_reference_test_decoder_runs_unmodified():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _staged("test/test_decoder.py"); _staged("test/hamming_7-4.csv")
    # cwd = the staging directory: the file opens "test/hamming_7-4.csv" by relative path
    r = subprocess.run([sys.executable, "-m", "unittest", "-v", "test.test_decoder"], cwd=STAGE, env=_env(),
                       capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stderr[-3000:])
    assert r.returncode == 0, r.stderr[-3000:]
    import re
    ran = re.search(r"Ran (\d+) tests", r.stderr)
    assert ran and int(ran.group(1)) == 10 and r.stderr.strip().splitlines()[-1] == "OK", r.stderr[-500:]
    # it really was this package (not the compiled reference in oracle/_ref) that the file imported
    probe = subprocess.run([sys.executable, "-c", "import qamreconciliation, galois; print(qamreconciliation.__file__); "
                            "print(galois.__file__)"], cwd=STAGE, env=_env(), capture_output=True, text=True)
    lines = probe.stdout.split()
    assert lines[0].startswith(PKG) and lines[1].startswith(STUBS), probe.stdout + probe.stderr


@pytest.mark.parametrize("flags,snr", [([], ("3.0", "6.0")), (["--hard"], ("4.0", "8.0")), (["--direct"], ("3.0", "6.0"))])
def test_reference_sim_script_runs_unmodified(tmp_path, flags, snr):
    """`python sims/sim_reconciliation.py EDGEFILE ...` of the reference, by path, unchanged: BER/FER fall from 1
    below the waterfall to 0 above it."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pandas as pd
    script = _staged("sims/sim_reconciliation.py")
    sys.path.insert(0, PKG)
    from qamreconciliation import codes
    vid, cid = codes.regular_ldpc(1296, 3, 6, seed=5)
    edge = tmp_path / "edges.csv"
    codes.write_edge_csv(str(edge), vid, cid)
    out = tmp_path / "out.csv"
    r = subprocess.run([sys.executable, script, str(edge), "--out", str(out), "--simloops", "300", "--nsnr", "3",
                        "--snr", snr[0], snr[1], "--ferr-count-min", "1000"] + flags, cwd=str(tmp_path), env=_env(),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    df = pd.read_csv(out)
    assert list(df.columns[1:]) == ["EsN0dB", "ber", "fer", "iters"] and len(df) == 3
    assert np.allclose(df.EsN0dB, np.linspace(float(snr[0]), float(snr[1]), 3))
    # (the script seeds nothing, like the reference: thresholds leave room for the Monte-Carlo spread of 300 frames)
    assert df.fer.iloc[0] > 0.5 and df.fer.iloc[-1] <= 0.01 and df.ber.iloc[-1] <= 1e-4
    assert df.ber.iloc[0] > df.ber.iloc[1] >= df.ber.iloc[2]
    assert 0 < df.iters.iloc[-1] < 15
