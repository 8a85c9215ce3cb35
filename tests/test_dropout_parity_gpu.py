"""GPU: numerical parity of the TIMED configuration — train mode, dropout 0.1 on every site (modeling.py:240, 296, 315, 355, 1007,
1018) — which the eval-mode golden tests cannot see.  The kernels draw their keep-decisions from a counter-based Philox stream
keyed by (seed, site, element); `vlpk_debug_dropout_mask` replays exactly those decisions, the test hands them to the fp32 oracle
through `oracle.vlp_oracle.MASK_PROVIDER`, and loss, logits and EVERY parameter gradient are compared at the tolerances of the
eval-mode tests.  A forward/backward mask mismatch in any fused kernel (attention bwd regenerating P's mask, the ReLU+dropout GEMM
epilogue, LayerNorm bwd) shows up here as a gradient error of order sqrt(p) ~ 30 %."""
import pytest
import torch

from oracle import vlp_oracle as O
from vlp_b200 import ops, synth
from vlp_b200 import vlp_modules as vm

from test_parity_gpu import TOL_HID, build, check_loss, compare_grads, make_config, rel, run_model

pytestmark = pytest.mark.gpu
P = 0.1


def _provider(seeds, dims, B):
    """oracle dropout site -> keep mask regenerated from the kernels' Philox streams."""
    L, H, heads, R = dims.seq_len, dims.hidden, dims.heads, dims.regions
    used = []

    def provide(site, shape):
        kind = site[0]
        if kind in ("vis_embed", "vis_pe_embed"):
            sid = (1 << 21) + (1 if kind == "vis_embed" else 2)
            m = ops.dropout_keep_mask(P, seeds[f"linear:{sid}"], sid, B * R * H).view(B, R, H)
        elif kind == "embed":
            m = ops.dropout_keep_mask(P, seeds["embed"], 1 << 20, B * L * H).view(B, L, H)
        elif kind == "attn":
            m = ops.dropout_keep_mask(P, seeds["encoder"], site[1] * 8 + 0, B * heads * L * 128).view(B, heads, L, 128)[..., :L]
        else:
            m = ops.dropout_keep_mask(P, seeds["encoder"], site[1] * 8 + (1 if kind == "hid1" else 2), B * L * H).view(B, L, H)
        assert tuple(m.shape) == tuple(shape), (site, m.shape, shape)
        frac = float(m.float().mean())
        assert abs(frac - (1 - P)) < 0.02, (site, frac)          # Bernoulli(0.9) keep rate
        used.append(site)
        return m.cpu().float()

    return provide, used


@pytest.mark.parametrize("dims,B,mode,tasks", [(synth.SMALL_L123, 4, "mix", "img2txt"), (synth.VlpDims(vocab=2000, layers=2), 3, "s2s", "img2txt"),
                                                (synth.SMALL_L123, 3, "bi", "vqa2")])
def test_training_mode_dropout_matches_oracle_with_replayed_masks(dims, B, mode, tasks):
    torch.manual_seed(1234)
    batch = synth.make_batch(dims, B, seed=77, mode=mode, ragged=True, tasks=tasks)
    model = build(dims, tasks, drop=P).train()
    ops.SEED_LOG = []
    try:
        losses = run_model(model, batch, tasks)
        sum(l.float().sum() for l in losses).backward()
        torch.cuda.synchronize()
        seeds = dict(ops.SEED_LOG)
    finally:
        ops.SEED_LOG = None
    assert set(seeds) == {"encoder", "embed", f"linear:{(1 << 21) + 1}", f"linear:{(1 << 21) + 2}"}, seeds

    sd = synth.make_state_dict(dims, 0, tasks)
    for k, v in sd.items():
        if k != "cls.predictions.decoder.weight":
            v.requires_grad_(True)
    provide, used = _provider(seeds, dims, B)
    O.MASK_PROVIDER = provide
    try:
        ref_losses, aux = O.pretraining_loss(sd, dims, batch, tasks=tasks, p_hidden=P, p_attn=P, training=True, return_all=True)
        sum(l.float().sum() for l in ref_losses).backward()
    finally:
        O.MASK_PROVIDER = None
    assert len(used) == 3 + 3 * dims.layers, used
    for got, ref in zip(losses, ref_losses):
        check_loss(got, ref)
    if tasks != "vqa2":
        assert rel(model.last_prediction_scores, aux["logits"]) < TOL_HID
    # the masks must actually matter: the same weights in eval mode give a visibly different loss
    ev = run_model(build(dims, tasks, drop=0.0).eval(), batch, tasks)
    assert abs(float(sum(l.float().sum() for l in ev)) - float(sum(l.float().sum() for l in losses))) > 1e-3
    ref_grads = {k: {"full": v.grad} for k, v in sd.items() if v.grad is not None}

    def drift():
        """the reference algorithm's own fp32 -> bf16 drift under the SAME masks (oracle re-run with bf16 weights and inputs)"""
        lo = {k: v.detach().to(torch.bfloat16) for k, v in synth.make_state_dict(dims, 0, tasks).items()}
        lo["cls.predictions.decoder.weight"] = lo["bert.embeddings.word_embeddings.weight"]
        for v in lo.values():
            v.requires_grad_(True)
        lb = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v) for k, v in batch.items()}
        O.MASK_PROVIDER = _provider(seeds, dims, B)[0]
        try:
            ll = O.pretraining_loss(lo, dims, lb, tasks=tasks, p_hidden=P, p_attn=P, training=True)
            sum(l.float().sum() for l in ll).backward()
        finally:
            O.MASK_PROVIDER = None
        return {k: rel(lo[k].grad, g["full"]) for k, g in ref_grads.items() if lo[k].grad is not None and float(g["full"].norm()) > 0}

    worst = compare_grads(model, ref_grads, drift_fn=drift)
    print(f"dropout parity {dims.hidden}H/{dims.layers}L B={B} {mode}/{tasks}: worst grad rel-L2 {worst:.3e}")
