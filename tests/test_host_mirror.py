"""C++ host-side mirror of the reference interface (consensus_b200/host/): api.Verifier, the batched
call-site restatements and the aggregator.  The C++ tests mirror TestBadCommit, TestNormalPath,
TestBadPrepare, TestValidateLastDecision, TestQuorum, TestReqPoolPrune,
TestControllerLeaderRequestHandling of /root/reference/internal/bft/*_test.go."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "consensus_b200", "host")


def _binary():
    subprocess.check_call(["make", "-s", "-C", HOST, "host_tests"])
    return os.path.join(HOST, "host_tests")


def test_host_mirror_with_mock_verifier():
    out = subprocess.run([_binary(), "cpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "0 failures" in out.stdout
    for name in ["TestBadCommit", "TestNormalPath", "TestValidateLastDecision", "TestQuorum", "TestReqPoolPrune"]:
        assert name in out.stdout


@pytest.mark.gpu
def test_host_mirror_on_the_engine_with_real_signatures():
    out = subprocess.run([_binary(), "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "TestGpuVerifierEndToEnd" in out.stdout and "0 failures" in out.stdout
