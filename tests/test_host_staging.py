"""CPU self-check of the host staging pool and the float64 -> float32 narrowing used by the host-memory decode path
(ldpc_3gpp_matlab_b200/csrc/host_staging.cpp): what nrldpc_decode64 runs on the doubles a MEX gateway hands over
(NRLDPCDecoder.m:262-265)."""
import json
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_pool_and_narrowing(tmp_path):
    exe = tmp_path / "host_staging_check"
    src = ROOT / "ldpc_3gpp_matlab_b200" / "csrc"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", f"-I{src}", str(ROOT / "tests" / "stubs" / "host_staging_check.cpp"),
                    str(src / "host_staging.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), str(1 << 22)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rec = json.loads(r.stdout.strip().splitlines()[-1])
    assert rec["bad"] == 0 and rec["threads"] >= 1
