"""The kernels' arithmetic without a GPU: pir_b200/csrc/pirb_device.cuh is compiled for the host
(tests/cpp/device_math_host_test.cpp) and its FP64 modular products, multiply-accumulate chains, shared-memory NTT
passes and Galois gather are checked against 128-bit integer arithmetic and the oracle's transforms."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_arithmetic_on_the_host():
    subprocess.check_call(["make", "-C", ROOT, "build/device_math_host_test"], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(ROOT, "build", "device_math_host_test")], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and "DEVICE_MATH_HOST_TEST_OK" in out.stdout, out.stdout + out.stderr


def test_ciphertext_multiplication_kernels_on_the_host():
    """tests/cpp/ctmul_host_test.cpp: the ciphertext-multiplication mode's host setup (auxiliary bases, base-conversion
    constants) and kernel bodies (pirb_behz.cuh), a whole upper dimension emulated launch by launch, against the oracle."""
    subprocess.check_call(["make", "-C", ROOT, "build/ctmul_host_test"], stdout=subprocess.DEVNULL)
    out = subprocess.run([os.path.join(ROOT, "build", "ctmul_host_test")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "CTMUL_HOST_TEST_OK" in out.stdout, out.stdout + out.stderr
