"""Algorithm-level models of the kernels that were written after round 1's GPU budget ran out (csrc/optim.cu, tables.cu, head.cu,
mask_pack_warp): each model transliterates the kernel's index arithmetic and work decomposition into numpy / torch on CPU —
chunk -> tensor binary search, entry -> row mapping, per-thread online log-sum-exp stripes, ballot words — and is checked against
the oracle / plain torch math.  This pins the DESIGN of those kernels (decomposition, edge handling); their CUDA code is covered by
the gated GPU tests (tests/test_zz_*_gpu.py)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import bertadam_oracle as bo

CHUNK = 4096


def _find_tensor(prefix, chunk):            # optim.cu find_tensor
    lo, hi = 0, len(prefix) - 1
    while hi - lo > 1:
        mid = (lo + hi) >> 1
        if prefix[mid] <= chunk:
            lo = mid
        else:
            hi = mid
    return lo


def test_bertadam_chunked_multi_tensor_model_matches_oracle():
    params, wds, grads = bo.case()
    h = bo.CASE_HYPER
    ps = [p.clone().flatten().numpy() for p in params]
    ms = [np.zeros_like(p) for p in ps]
    vs = [np.zeros_like(p) for p in ps]
    rp = [p.clone() for p in params]
    rm = [torch.zeros_like(p) for p in params]
    rv = [torch.zeros_like(p) for p in params]
    n = [p.size for p in ps]
    prefix = np.concatenate(([0], np.cumsum([(k + CHUNK - 1) // CHUNK for k in n]))).astype(np.int32)
    n_chunks = int(prefix[-1])
    f32 = np.float32
    for t in range(bo.CASE_STEPS):
        gs = [g.flatten().numpy() for g in grads[t]]
        lr = f32(bo.lr_at(t, h["lr"], h["warmup"], h["t_total"], h["schedule"]))
        b1, omb1, b2, omb2, eps = f32(h["b1"]), f32(1.0 - h["b1"]), f32(h["b2"]), f32(1.0 - h["b2"]), f32(h["e"])
        sq = np.zeros(len(ps), dtype=f32)
        for c in range(n_chunks):                                   # adam_sqnorm_kernel: one partial per chunk
            ti = _find_tensor(prefix, c)
            base = (c - prefix[ti]) * CHUNK
            seg = gs[ti][base:base + CHUNK]
            sq[ti] += f32(np.sum(seg.astype(np.float64) ** 2))
        covered = [0] * len(ps)
        for c in range(n_chunks):                                   # adam_update_kernel
            ti = _find_tensor(prefix, c)
            base = (c - prefix[ti]) * CHUNK
            cnt = min(CHUNK, n[ti] - base)
            covered[ti] += cnt
            sl = slice(base, base + cnt)
            cc = f32(h["max_grad_norm"]) / (np.sqrt(sq[ti]) + f32(1e-6))
            coef = cc if cc < 1 else f32(1.0)
            g = gs[ti][sl] * coef
            ms[ti][sl] = omb1 * g + ms[ti][sl] * b1
            vs[ti][sl] = (omb2 * g) * g + vs[ti][sl] * b2
            u = ms[ti][sl] / (np.sqrt(vs[ti][sl]) + eps)
            if wds[ti] > 0:
                u = f32(wds[ti]) * ps[ti][sl] + u
            ps[ti][sl] = ps[ti][sl] - lr * u
        assert covered == n                                          # every element exactly once, ragged last chunks included
        for i in range(len(ps)):
            bo.step(rp[i], grads[t][i].clone(), rm[i], rv[i], t, weight_decay=wds[i], **h)
            for mine, ref in ((ps[i], rp[i]), (ms[i], rm[i]), (vs[i], rv[i])):
                ref = ref.flatten().numpy()
                assert np.abs(mine - ref).max() <= 2e-6 * np.abs(ref).max() + 1e-30, (t, i)


def _table_row(e, L, R, vis):               # tables.cu table_row
    n_tab = L - R if vis else L
    b, k = divmod(e, n_tab)
    l = (0 if k == 0 else R + k) if vis else k
    return b * L + l


def test_embedding_table_gradient_model_matches_torch_scatter():
    gen = torch.Generator().manual_seed(0)
    for (B, L, R, H, V, vis) in ((3, 15, 4, 16, 40, True), (2, 9, 0, 8, 12, False), (4, 123, 100, 8, 300, True)):
        P, T = 130, 6
        ids = torch.randint(0, V, (B, L), generator=gen)
        ids[:, 0] = 1
        tt = torch.randint(0, T, (B, L), generator=gen)
        dz = torch.randn(B * L, H, generator=gen).bfloat16().float()
        # model: phases of word_pos_kernel over the looked-up rows only + segmented type sums per 256-row slab
        n_entries = B * ((L - R) if vis else L)
        rows = [_table_row(e, L, R, vis) for e in range(n_entries)]
        keep = sorted(rows)
        expect_rows = sorted(b * L + l for b in range(B) for l in range(L) if (not vis) or l == 0 or l > R)
        assert keep == expect_rows                                   # exactly the rows that read the word / position tables
        scratch = torch.full((V, H), float("nan"))                   # uninitialised: only touched rows may be read
        for r in rows:
            scratch[ids.view(-1)[r]] = 0.0
        d_pos = torch.zeros(P, H)
        for r in rows:
            scratch[ids.view(-1)[r]] += dz[r]
            d_pos[r % L] += dz[r]
        d_word = torch.zeros(V, H, dtype=torch.bfloat16)
        for r in rows:
            d_word[ids.view(-1)[r]] = scratch[ids.view(-1)[r]].bfloat16()
        d_type = torch.zeros(T, H)
        for r0 in range(0, B * L, 256):
            slab = slice(r0, min(B * L, r0 + 256))
            for t in range(T):
                d_type[t] += (dz[slab] * (tt.view(-1)[slab] == t).unsqueeze(1)).sum(0)
        # reference: the torch scatter of ops.EmbedFn.backward
        ref_word = torch.zeros(V, H).index_add_(0, ids.view(-1)[keep], dz[keep])
        pos_idx = torch.tensor([r % L for r in keep])
        ref_pos = torch.zeros(P, H).index_add_(0, pos_idx, dz[keep])
        ref_type = torch.zeros(T, H).index_add_(0, tt.view(-1), dz)
        assert torch.allclose(d_word.float(), ref_word.bfloat16().float(), atol=2e-2, rtol=1e-2)
        assert torch.allclose(d_pos, ref_pos, atol=1e-4) and torch.allclose(d_type, ref_type, atol=1e-4)
        assert not torch.isnan(d_word.float()).any()


def _online_merge(m, s, m2, s2):            # head.cu online_merge
    mn = max(m, m2)
    return mn, s * math.exp(m - mn) + s2 * math.exp(m2 - mn)


def test_decoder_ce_row_model_matches_torch_cross_entropy():
    gen = torch.Generator().manual_seed(1)
    THREADS = 256
    for V in (1003, 29, 8 * THREADS + 5):
        Vp = (V + 7) // 8 * 8
        R = 5
        logits = torch.zeros(R, Vp)
        logits[:, :V] = (torch.randn(R, V, generator=gen) * 3).bfloat16().float()
        labels = torch.randint(0, V, (R,), generator=gen)
        labels[2] = -100
        dloss = torch.rand(R, generator=gen)
        lse, loss = torch.zeros(R), torch.zeros(R)
        for r in range(R):
            parts = []
            for tid in range(THREADS):                               # per-thread stripes of 8 columns, stride 8 * THREADS
                m, s = -3.0e38, 0.0
                for c in range(tid * 8, V, THREADS * 8):
                    x = [float(logits[r, c + j]) for j in range(8) if c + j < V]
                    cm = max(x)
                    cs = sum(math.exp(v - cm) for v in x)
                    m, s = _online_merge(m, s, cm, cs)
                parts.append((m, s))
            M, S = parts[0]
            for m2, s2 in parts[1:]:
                M, S = _online_merge(M, S, m2, s2)
            lse[r] = M + math.log(S)
            y = int(labels[r])
            loss[r] = lse[r] - logits[r, y] if 0 <= y < V else 0.0
        ref = F.cross_entropy(logits[:, :V], labels, reduction="none", ignore_index=-100)
        assert torch.allclose(loss, ref, atol=1e-4)
        # backward model: (exp(x - lse) - onehot) * dloss, zero in pad columns and for the ignored row
        live = (labels >= 0) & (labels < V)
        d = torch.exp(logits - lse[:, None])
        d[:, V:] = 0
        d[torch.arange(R)[live], labels[live]] -= 1
        d = d * (dloss * live)[:, None]
        x = logits[:, :V].clone().requires_grad_(True)
        (F.cross_entropy(x, labels, reduction="none", ignore_index=-100) * dloss).sum().backward()
        assert torch.allclose(d[:, :V], x.grad, atol=1e-5) and float(d[:, V:].abs().sum()) == 0 and float(d[2].abs().sum()) == 0


def test_warp_ballot_mask_pack_model_matches_row_walk():
    gen = torch.Generator().manual_seed(3)
    for kv in (15, 77, 123, 128):
        row = (torch.rand(kv, generator=gen) < 0.5).tolist()
        walk = [0, 0, 0, 0]
        for j in range(kv):
            if row[j]:
                walk[j >> 5] |= 1 << (j & 31)
        ballot = []
        for i in range(4):                                           # word i = ballot over lanes of element lane + 32 i
            w = 0
            for lane in range(32):
                j = lane + 32 * i
                if j < kv and row[j]:
                    w |= 1 << lane
            ballot.append(w)
        assert ballot == walk
