"""Kernel-level parity (GPU): each kernel family of libvlpk.so, called through the C ABI, against plain PyTorch fp32 math of
the same op at small and production tile shapes (cases live in tools/bringup.py so they can also run as a harness)."""
import pytest

from tools import bringup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["gemm_kk", "gemm_epi", "gemm_dgrad", "gemm_wgrad", "attn", "attn_full", "attn_common_mode", "rowops"])
def test_kernel_family(case):
    assert bringup.CASES[case]()
