"""GPU: the finer-grained C-ABI entry points (one fwd and one bwd symbol per fused op) agree with the composite
calls the module surface uses.  Call sequences live in tools/abi_cases.py (also dry-run on CPU by test_modules_cpu.py)."""
import pytest
import torch

from tools import abi_cases

pytestmark = pytest.mark.gpu


def test_ffn_bwd_plus_mha_bwd_equals_layer_bwd():
    o = abi_cases.split_backward_case("cuda")
    torch.cuda.synchronize()
    assert torch.isfinite(o["y"].float()).all() and float(o["dx_a"].float().abs().sum()) > 0
    assert torch.equal(o["dx_a"], o["dx_b"])                      # bf16 epilogue outputs: same kernels, same inputs
    a, b = o["arena_a"].double(), o["arena_b"].double()           # fp32 split-K reduce-add: summation order may differ
    assert float((a - b).norm() / a.norm()) < 1e-5
    assert float(a.abs().max()) > 0


def test_mha_incr_fwd_equals_mha_fwd_with_history():
    o = abi_cases.incremental_case("cuda")
    torch.cuda.synchronize()
    assert torch.isfinite(o["y1_mha"].float()).all()
    assert torch.equal(o["y1_mha"], o["y1_incr"])
