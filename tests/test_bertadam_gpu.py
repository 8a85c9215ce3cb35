"""GPU parity of the fused BertAdam step (csrc/optim.cu through vlpk_bertadam_step) against the oracle restatement of
pytorch_pretrained_bert/optimization.py:112-182 and against the reference's own outputs in tests/golden/bertadam.pt.

Tolerance: fp32 arithmetic on both sides; differences come from FMA contraction and the fp32-vs-double clip coefficient, i.e. a
few ulp of each tensor's scale (sums of opposite-signed terms cancel, so the bound is relative to the tensor's max, 2e-6)."""
import os

import pytest
import torch

from oracle import bertadam_oracle as bo
from vlp_b200 import optimization as opt_mod

pytestmark = pytest.mark.gpu


def _close(x, y, what, tol=2e-6):
    x, y = x.detach().float().cpu(), y.detach().float().cpu()
    err, scale = float((x - y).abs().max()), float(y.abs().max())
    assert err <= tol * scale + 1e-30, (what, err, scale)


def _groups(ps, wds):
    return [{"params": [p for p, w in zip(ps, wds) if w > 0], "weight_decay": 0.01},
            {"params": [p for p, w in zip(ps, wds) if w == 0], "weight_decay": 0.0}]


def test_fp32_parameters_match_reference_golden(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "bertadam.pt"))
    params, wds, grads = bo.case()
    ps = [torch.nn.Parameter(p.clone().cuda()) for p in params]
    opt = opt_mod.BertAdam(_groups(ps, wds), **bo.CASE_HYPER)
    for t in range(bo.CASE_STEPS):
        for p, g in zip(ps, grads[t]):
            p.grad = g.clone().cuda()
        opt.step()
        torch.cuda.synchronize()
        for i, p in enumerate(ps):
            _close(p, gold["steps"][t]["p"][i], ("p", t, i))
            _close(opt.state[p]["next_m"], gold["steps"][t]["m"][i], ("m", t, i))
            _close(opt.state[p]["next_v"], gold["steps"][t]["v"][i], ("v", t, i))
            assert torch.equal(p.grad.cpu(), grads[t][i])          # gradients are not rescaled in place (documented difference)
            assert opt.state[p]["step"] == t + 1 and "master" not in opt.state[p]


def test_bf16_parameters_follow_an_fp32_master_copy():
    params, wds, grads = bo.case()
    ps = [torch.nn.Parameter(p.clone().bfloat16().cuda()) for p in params]
    # oracle: fp32 arithmetic from the bf16-rounded start, bf16-rounded gradients
    rp = [p.detach().float().cpu() for p in ps]
    rm = [torch.zeros_like(p) for p in rp]
    rv = [torch.zeros_like(p) for p in rp]
    opt = opt_mod.BertAdam(_groups(ps, wds), **bo.CASE_HYPER)
    for t in range(bo.CASE_STEPS):
        gs = [g.bfloat16() for g in grads[t]]
        for p, g in zip(ps, gs):
            p.grad = g.clone().cuda()
        opt.step()
        torch.cuda.synchronize()
        for i in range(len(ps)):
            bo.step(rp[i], gs[i].float(), rm[i], rv[i], t, weight_decay=wds[i], **bo.CASE_HYPER)
            st = opt.state[ps[i]]
            _close(st["master"], rp[i], ("master", t, i))
            _close(st["next_m"], rm[i], ("m", t, i))
            _close(st["next_v"], rv[i], ("v", t, i), tol=6e-6)     # v ~ clip^2: twice the relative error of the fp32-vs-double clip factor
            assert torch.equal(ps[i].detach(), st["master"].bfloat16())    # the bf16 parameter is the rounding of its master copy


def test_no_clipping_constant_lr_and_skipped_parameters():
    gen = torch.Generator().manual_seed(5)
    w = torch.nn.Parameter((torch.randn(1000, 33, generator=gen) * 0.1).cuda())
    frozen = torch.nn.Parameter(torch.randn(10, generator=gen).cuda())           # never gets a gradient
    opt = opt_mod.BertAdam([w, frozen], lr=1e-2, max_grad_norm=-1, weight_decay=0.0)
    rp, rm, rv = w.detach().cpu().clone(), torch.zeros(1000, 33), torch.zeros(1000, 33)
    before = frozen.detach().clone()
    for t in range(2):
        g = torch.randn(1000, 33, generator=gen) * 5.0                            # ||g|| >> 1 but clipping is off
        w.grad = g.clone().cuda()
        opt.step()
        bo.step(rp, g.clone(), rm, rv, t, lr=1e-2, max_grad_norm=-1, weight_decay=0.0)
    torch.cuda.synchronize()
    _close(w, rp, "p")
    assert torch.equal(frozen.detach(), before) and len(opt.state[frozen]) == 0
