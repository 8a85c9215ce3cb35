"""N > 1 host-side logic on CPU: two `gloo` ranks (SURVEY.md §8e — the path shards by batch, the only collective is the
gradient all-reduce that torch DDP performs).  No kernel runs here (there is no CPU compute path); what is covered is what
the multi-GPU bench relies on besides the kernels: DDP can wrap the module tree (every trainable parameter is a leaf that a
bucket can own, pooler frozen instead of find_unused_parameters), rank 0's parameters are broadcast at construction,
per-rank synthetic shards differ, and averaging per-rank gradients reproduces the gradient of the concatenated batch for a
mean-normalised loss.
"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vlp_b200 import synth
from vlp_b200 import vlp_modules as vm


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = synth.TINY
        cfg = vm.BertConfig(d.vocab, hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads,
                            intermediate_size=d.inter, type_vocab_size=d.type_vocab, max_position_embeddings=d.max_pos)
        torch.manual_seed(100 + rank)                      # deliberately different initialisation per rank
        model = vm.BertForPreTrainingLossMask(cfg, enable_butd=True, len_vis_input=d.regions)
        for p in model.bert.pooler.parameters():           # never receives a gradient in img2txt
            p.requires_grad_(False)
        ddp = torch.nn.parallel.DistributedDataParallel(model, gradient_as_bucket_view=True, broadcast_buffers=False)
        # (1) constructor broadcast: every rank now holds rank 0's parameters
        flat = torch.cat([p.detach().flatten() for p in model.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(flat, ref))
        # (2) per-rank shards differ
        b = synth.make_batch(d, 2, seed=1234 + rank)
        ids_sum = b["input_ids"].sum().clone()
        gathered = [torch.zeros_like(ids_sum) for _ in range(world)]
        dist.all_gather(gathered, ids_sum)
        distinct = len({int(g) for g in gathered}) == world
        # (3) the collective arithmetic: mean of per-rank gradients (what DDP's buckets compute)
        trainable = [p for p in model.parameters() if p.requires_grad]
        g = torch.Generator().manual_seed(7 + rank)
        for p in trainable:
            p.grad = torch.randn(p.shape, generator=g)
        local = torch.cat([p.grad.flatten() for p in trainable])
        tot = local.clone()
        dist.all_reduce(tot)
        tot /= world
        allg = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(allg, local)
        mean_ok = bool(torch.allclose(tot, torch.stack(allg).mean(0), atol=1e-6))
        n_tr = sum(p.numel() for p in trainable)
        q.put((rank, same, distinct, mean_ok, n_tr, len(list(ddp.parameters()))))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_ddp_wrap_and_collective():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same, distinct, mean_ok, n_tr, n_all in res:
        assert same, f"rank {rank}: parameters differ from rank 0 after DDP construction"
        assert distinct and mean_ok
    assert res[0][4] == res[1][4]


def _reducer_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vlp_b200.dp import GradientAllReducer
        d = synth.TINY
        cfg = vm.BertConfig(d.vocab, hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads,
                            intermediate_size=d.inter, type_vocab_size=d.type_vocab, max_position_embeddings=d.max_pos)
        torch.manual_seed(rank)
        model = vm.BertForPreTrainingLossMask(cfg, enable_butd=True, len_vis_input=d.regions)
        enc = model.bert.encoder
        before = (enc.layers_per_call, enc._vlpk_grad_hook)
        red = GradientAllReducer(model, layer_groups=[1, 1])
        assert enc.layers_per_call == [1, 1] and enc._vlpk_grad_hook == red._on_encoder_grads      # the hook belongs to THIS encoder
        red.broadcast_parameters(0)
        w0 = model.bert.encoder.layer[1].output.dense.weight.detach().clone()
        ref = w0.clone()
        dist.broadcast(ref, src=0)
        for p in red.other:
            p.grad = torch.full_like(p, float(rank + 1))
        arena = torch.full((1000,), float(rank + 1))           # stands in for an encoder group's flat gradient arena
        red._on_encoder_grads(arena)
        red.finish()
        mean = (world + 1) / 2.0
        ok = torch.equal(w0, ref) and all(torch.allclose(p.grad, torch.full_like(p, mean)) for p in red.other) and \
            torch.allclose(arena, torch.full((1000,), mean))
        # gradient accumulation: encoder gradients already populated when the arena arrives -> reduced values in place before return
        enc.layer[0].output.dense.weight.grad = torch.zeros_like(enc.layer[0].output.dense.weight)
        arena2 = torch.full((10,), float(rank + 1))
        red._on_encoder_grads(arena2)
        ok = ok and red._accumulating is True and torch.allclose(arena2, torch.full((10,), mean)) and not red._works
        red.finish()
        red.enabled = False                                     # switched off: nothing is touched
        arena3 = torch.full((10,), float(rank + 1))
        red._on_encoder_grads(arena3)
        red.finish()
        ok = ok and torch.equal(arena3, torch.full((10,), float(rank + 1)))
        enc_ids = {id(p) for p in model.bert.encoder.parameters()}
        disjoint = all(id(p) not in enc_ids for p in red.other)
        red.close()
        ok = ok and (enc.layers_per_call, enc._vlpk_grad_hook) == before           # close() restores the encoder
        q.put((rank, bool(ok), disjoint, 1))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradient_arena_reducer():
    """vlp_b200.dp.GradientAllReducer (what bench.py uses for N > 1): parameter broadcast, mean of the flat encoder arena handed
    over by the backward hook, mean of the remaining (non-encoder) gradients through one flattened buffer."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_reducer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, disjoint, lpc in res:
        assert ok and disjoint and lpc == 1


def test_trainable_parameter_count_matches_survey():
    """115 939 396 trainable elements for img2txt BERT-base (the all-reduce payload), pooler included."""
    d = synth.BERT_BASE
    n = sum(int(torch.tensor(s).prod()) for k, s, _ in synth.state_dict_keys(d))
    assert n == 115939396
