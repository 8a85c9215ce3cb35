"""Data-parallel gradient all-reduce for the VLP hot path (the path shards by batch; the one collective
is the mean of all gradients per step — `DistributedDataParallel` in the reference, vlp/run_img2txt_dist.py:386).

torch DDP works unchanged on vlp_b200 modules.  This module is the B200-first alternative used by bench.py: the fused
encoder backward already produces each layer group's gradients as ONE contiguous bf16 arena, so the arena itself is handed to
NCCL (`all_reduce`, AVG, asynchronously, the moment the group's backward finishes) while earlier groups are still computing —
no per-parameter bucket copies, no autograd hooks, no graph walk for unused parameters.  The few remaining parameters
(embeddings, region projections, heads) are reduced as one flattened buffer after backward.
"""
import os

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops


class GradientAllReducer:
    def __init__(self, model, group=None, layers_per_call=3, reserved_sms=None):
        """reserved_sms (experiment, default 0 / env VLP_DP_RESERVED_SMS): while an arena all-reduce is in flight the persistent
        GEMM grids launched after it leave that many SMs to NCCL's CTAs (vlpk_set_reserved_sms); restored by finish()."""
        self.reserved_sms = int(os.environ.get("VLP_DP_RESERVED_SMS", "0")) if reserved_sms is None else int(reserved_sms)
        self.group = group
        self.world = dist.get_world_size(group)
        self.model = model
        self.backend = dist.get_backend(group)
        enc = model.bert.encoder
        groups = os.environ.get("VLP_DP_GROUPS")          # experiment: explicit group sizes from layer 0 up, e.g. "1,2,3,3,3"
        enc.layers_per_call = [int(k) for k in groups.split(",")] if groups else layers_per_call
        enc_ids = {id(p) for p in enc.parameters()}
        self.other = [p for p in model.parameters() if p.requires_grad and id(p) not in enc_ids]
        ops.set_encoder_grad_hook(self._on_encoder_grads)
        self._works = []

    def broadcast_parameters(self, src=0):
        for p in self.model.parameters():
            dist.broadcast(p.data, src=src, group=self.group)

    def _reduce(self, t, async_op):
        if self.backend == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=False)   # gloo: no AVG
        t.div_(self.world)
        return w if async_op else None

    def _on_encoder_grads(self, arena):
        """Called by EncoderStackFn.backward with the flat gradient arena of one layer group (views of it become .grad)."""
        w = self._reduce(arena, async_op=True)
        if w is not None:
            self._works.append(w)
        if self.reserved_sms > 0:
            L.lib().vlpk_set_reserved_sms(self.reserved_sms)

    def finish(self):
        """After loss.backward(): reduce the non-encoder gradients and wait for everything in flight."""
        grads = [p.grad for p in self.other if p.grad is not None]
        if grads:
            flat = torch._utils._flatten_dense_tensors(grads)
            self._reduce(flat, async_op=False)
            for g, r in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                g.copy_(r)
        for w in self._works:
            w.wait()
        self._works.clear()
        if self.reserved_sms > 0:
            L.lib().vlpk_set_reserved_sms(0)

    def close(self):
        ops.set_encoder_grad_hook(None)
