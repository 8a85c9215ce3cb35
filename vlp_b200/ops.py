"""torch.autograd glue between the nn.Module surface (vlp_modules.py) and the C ABI (libvlpk.so).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape.  Every forward/backward
body is one (or a few) calls into the hand-written CUDA library — there is no PyTorch implementation of
these ops anywhere in the package, so a missing library is a hard error.
"""
import ctypes as C
import itertools
import os

import torch

from . import _lib as L

BF16 = torch.bfloat16

# ------------------------------------------------------------------------------------------------
# dropout seeding
# ------------------------------------------------------------------------------------------------
_seed_counter = itertools.count(1)
_seed_dev = None  # optional int64 CUDA tensor: added to the host seed on device (CUDA-graph replays)


def set_device_seed_tensor(t):
    """Register a 1-element int64 CUDA tensor whose value is added to every dropout seed at kernel run time.
    Increment it between CUDA-graph replays to get fresh masks from a frozen launch sequence."""
    global _seed_dev
    _seed_dev = t


SEED_LOG = None  # tests: set to a list to record (kind, seed) of every dropout stream drawn (kind: "encoder", "linear:<site>", "embed")


def next_seed(kind=None):
    seed = (torch.initial_seed() * 1000003 + next(_seed_counter) * 7919) & 0x7FFFFFFFFFFFFFFF
    if SEED_LOG is not None:
        SEED_LOG.append((kind, seed))
    return seed


def dropout_keep_mask(p, seed, site, n):
    """uint8 [n] keep decisions of dropout site `site` under `seed` (vlpk_debug_dropout_mask) — test support."""
    n8 = (n + 7) // 8 * 8
    out = torch.empty(n8, dtype=torch.uint8, device="cuda")
    L.call("vlpk_debug_dropout_mask", _drop(p, seed), site, n8, out.data_ptr(), L.stream())
    return out[:n]


def _drop(p, seed):
    if p <= 0.0 or seed is None:
        return None
    return L.VlpkDropout(float(p), int(seed), None if _seed_dev is None else _seed_dev.data_ptr())


def _bf16c(t):
    """bf16 + contiguous view/copy of a tensor (parameters of a bf16 model pass through untouched)."""
    if t is None:
        return None
    if t.dtype != BF16:
        t = t.to(BF16)
    return t if t.is_contiguous() else t.contiguous()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"vlp_b200: {what} must live on a CUDA device (no CPU path exists)")


# ------------------------------------------------------------------------------------------------
# attention mask -> bitmask
# ------------------------------------------------------------------------------------------------
def pack_mask(mask, mode="additive"):
    """[B,1,R,KV] / [B,R,KV] additive (0/-10000) or 0/1 mask -> int32 [B,R,4] 'attend' bitmask (R may be 1)."""
    _require_cuda(mask, "attention mask")
    if mask.dim() == 4:
        mask = mask[:, 0]
    if mask.dim() == 2:
        mask = mask[:, None, :]
    mask = mask.contiguous()
    B, R, KV = mask.shape
    dt = {torch.float32: 1, torch.bfloat16: 0, torch.int64: 2}.get(mask.dtype)
    if dt is None:
        mask = mask.float()
        dt = 1
    out = torch.empty(B, R, 4, device=mask.device, dtype=torch.int32)
    L.call("vlpk_mask_pack", mask.data_ptr(), dt, 0 if mode == "additive" else 1, B, R, KV, R * KV, KV, out.data_ptr(), L.stream())
    return out


# ------------------------------------------------------------------------------------------------
# BertLayer stack
# ------------------------------------------------------------------------------------------------
PARAMS_PER_LAYER = 16  # order == _lib.WEIGHT_FIELDS


def _layer_sizes(H, I):
    # fp32 gradient arena layout of one layer, order == _lib.GRAD_FIELDS
    return [3 * H * H, 3 * H, H * H, H, H, H, I * H, I, H * I, H, H, H]


def _grad_views(arena, H, I):
    """Views of one layer's fp32 (or converted) arena in the order of WEIGHT_FIELDS."""
    sizes = _layer_sizes(H, I)
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    seg = {n: arena[offs[i]:offs[i + 1]] for i, n in enumerate(L.GRAD_FIELDS)}
    wqkv, bqkv = seg["wqkv"].view(3, H, H), seg["bqkv"].view(3, H)
    return [wqkv[0], wqkv[1], wqkv[2], bqkv[0], bqkv[1], bqkv[2], seg["wo"].view(H, H), seg["bo"], seg["ln1_g"], seg["ln1_b"],
            seg["w1"].view(I, H), seg["b1"], seg["w2"].view(H, I), seg["b2"], seg["ln2_g"], seg["ln2_b"]]


class _Acts:
    """Per-layer activation buffers (one bf16 + one fp32 allocation for the whole stack)."""

    def __init__(self, n_layers, B, Lq, H, heads, I, device, Lkv=None, drop_bits=False):
        M = B * Lq
        self.bf_sizes = [("qkv", M * 3 * H), ("ctx", M * H), ("t1", M * H), ("y1", M * H), ("u", M * I), ("hmid", M * I), ("t2", M * H),
                         ("y", M * H)]
        if Lkv is not None:
            self.bf_sizes.append(("kv", B * Lkv * 2 * H))
        # lse rounded up to 4 floats so that the float2 statistics behind it stay 8-byte aligned for odd B * heads * Lq
        self.f_sizes = [("lse", (B * heads * Lq + 3) // 4 * 4), ("stats1", 2 * M), ("stats2", 2 * M)]
        per_bf = sum(s for _, s in self.bf_sizes)
        per_f = sum(s for _, s in self.f_sizes)
        self.bf = torch.empty(n_layers, per_bf, device=device, dtype=BF16)
        self.f32 = torch.empty(n_layers, per_f, device=device, dtype=torch.float32)
        self.structs = (L.VlpkLayerActs * n_layers)()
        # training with dropout: 1 bit per attention probability (128 key slots per query row), written by the forward attention
        # kernel and re-read by the backward one instead of re-evaluating Philox
        self.bits = torch.empty(n_layers, B * heads * Lq * 16, device=device, dtype=torch.uint8) if drop_bits else None
        self.y = []
        for i in range(n_layers):
            st = self.structs[i]
            off = 0
            base = self.bf[i]
            for name, sz in self.bf_sizes:
                setattr(st, name, base[off:off + sz].data_ptr())
                if name == "y":
                    self.y.append(base[off:off + sz].view(B, Lq, H))
                off += sz
            if Lkv is None:
                st.kv = None
            st.drop_attn = None
            if self.bits is not None:
                st.drop_attn = self.bits[i].data_ptr()
            off = 0
            basef = self.f32[i]
            for name, sz in self.f_sizes:
                setattr(st, name, basef[off:off + sz].data_ptr())
                off += sz


def _weight_structs(params, n_layers):
    ws = (L.VlpkLayerWeights * n_layers)()
    for i in range(n_layers):
        for j, name in enumerate(L.WEIGHT_FIELDS):
            setattr(ws[i], name, params[i * PARAMS_PER_LAYER + j].data_ptr())
    return ws


class EncoderStackFn(torch.autograd.Function):
    """n_layers x BertLayer (modeling.py:367-402) in one C call each way.  Returns every layer's output."""

    @staticmethod
    def forward(ctx, hidden, mask_bits, cfg, *params):
        n_layers, heads, I, p_attn, p_hidden, training = cfg[:6]
        _require_cuda(hidden, "hidden_states")
        ctx.set_materialize_grads(False)   # unused layer outputs must arrive as None in backward, not as zero tensors
        x = _bf16c(hidden)
        B, Lq, H = x.shape
        pk = [_bf16c(p) for p in params]
        seed = next_seed("encoder") if (training and (p_attn > 0 or p_hidden > 0)) else None
        acts = _Acts(n_layers, B, Lq, H, heads, I, x.device, drop_bits=seed is not None)
        shape = L.VlpkShape(B, Lq, Lq, H, heads, I)
        ws = _weight_structs(pk, n_layers)
        drop = _drop(max(p_attn, p_hidden), seed)
        L.call("vlpk_encoder_fwd", C.byref(shape), n_layers, ws, x.data_ptr(), mask_bits.data_ptr(), mask_bits.shape[1], acts.structs,
               float(p_attn if training else 0.0), float(p_hidden if training else 0.0), drop, L.stream())
        ctx.cfg = cfg
        ctx.seed = seed
        ctx.acts = acts
        ctx.x = x
        ctx.mask_bits = mask_bits
        ctx.pk = pk
        ctx.param_dtypes = [p.dtype for p in params]
        ctx.mark_non_differentiable(mask_bits)
        outs = tuple(acts.y)
        acts.y = None        # ctx keeps the buffers (acts.bf / raw pointers), not the returned tensor objects: no output -> grad_fn -> ctx ->
        return outs          # output reference cycle, so an un-backpropagated training forward is freed by reference counting

    @staticmethod
    def backward(ctx, *dys):
        n_layers, heads, I, p_attn, p_hidden, training = ctx.cfg[:6]
        grad_hook = ctx.cfg[6] if len(ctx.cfg) > 6 else None      # data parallelism: callable(flat gradient arena) of the owning BertEncoder
        if ctx.acts is None:
            raise RuntimeError("vlp_b200: the fused BertLayer stack frees its activations in backward; a second backward "
                               "(retain_graph=True) is not supported")
        x, acts = ctx.x, ctx.acts
        B, Lq, H = x.shape
        M = B * Lq
        dev = x.device
        if dys[-1] is None:
            dys = list(dys)
            dys[-1] = torch.zeros_like(x)
        dyc = [None if d is None else _bf16c(d) for d in dys]
        dy_ptrs = (C.c_void_p * n_layers)(*[None if d is None else d.data_ptr() for d in dyc])
        per_layer = sum(_layer_sizes(H, I))
        arena = torch.zeros(n_layers, per_layer, device=dev, dtype=torch.float32)
        gs = (L.VlpkLayerGrads * n_layers)()
        sizes = _layer_sizes(H, I)
        for i in range(n_layers):
            off = 0
            for name, sz in zip(L.GRAD_FIELDS, sizes):
                setattr(gs[i], name, arena[i, off:off + sz].data_ptr())
                off += sz
        scr_sizes = {"dz2": M * H, "dt2": M * H, "du": M * I, "dy1": M * H, "dz1": M * H, "dt1": M * H, "dctx": M * H, "dqkv": 3 * M * H,
                     "dx": M * H}
        scratch = torch.empty(sum(scr_sizes.values()), device=dev, dtype=BF16)
        ws_s = L.VlpkBwdScratch()
        off = 0
        for name in L.SCRATCH_FIELDS:
            setattr(ws_s, name, scratch[off:off + scr_sizes[name]].data_ptr())
            off += scr_sizes[name]
        dx0 = torch.empty_like(x)
        shape = L.VlpkShape(B, Lq, Lq, H, heads, I)
        ws = _weight_structs(ctx.pk, n_layers)
        drop = _drop(max(p_attn, p_hidden), ctx.seed)
        L.call("vlpk_encoder_bwd", C.byref(shape), n_layers, ws, x.data_ptr(), ctx.mask_bits.data_ptr(), ctx.mask_bits.shape[1], acts.structs,
               dy_ptrs, dx0.data_ptr(), gs, C.byref(ws_s), float(p_attn if training else 0.0), float(p_hidden if training else 0.0), drop,
               L.stream())
        # gradient arena -> parameter dtype (one conversion kernel for the whole stack)
        if all(dt == BF16 for dt in ctx.param_dtypes):
            garena = torch.empty(n_layers, per_layer, device=dev, dtype=BF16)
            L.call("vlpk_f32_to_bf16", arena.data_ptr(), garena.data_ptr(), arena.numel(), L.stream())
        else:
            garena = arena
        if grad_hook is not None:
            grad_hook(garena)            # e.g. asynchronous NCCL all-reduce of this group's gradients (dp.GradientAllReducer)
        grads = []
        for i in range(n_layers):
            for v, dt in zip(_grad_views(garena[i], H, I), ctx.param_dtypes[i * PARAMS_PER_LAYER:(i + 1) * PARAMS_PER_LAYER]):
                grads.append(v if v.dtype == dt else v.to(dt))
        ctx.acts = None
        return (dx0, None, None) + tuple(grads)


def layer_incremental_fwd(hidden, history, mask_bits, heads, I, params):
    """BertLayer.forward with history_states (modeling.py:273-277, 389-390): inference only, q rows = hidden,
    kv rows = cat(history, hidden)."""
    x = _bf16c(hidden)
    xkv = _bf16c(torch.cat((history.to(x.dtype), x), dim=1))
    B, Lq, H = x.shape
    Lkv = xkv.shape[1]
    pk = [_bf16c(p) for p in params]
    acts = _Acts(1, B, Lq, H, heads, I, x.device, Lkv=Lkv)
    shape = L.VlpkShape(B, Lq, Lkv, H, heads, I)
    ws = _weight_structs(pk, 1)
    L.call("vlpk_layer_fwd", C.byref(shape), ws, x.data_ptr(), xkv.data_ptr(), mask_bits.data_ptr(), mask_bits.shape[1], acts.structs, 0.0, 0.0,
           None, 0, L.stream())
    return acts.y[0]


# ------------------------------------------------------------------------------------------------
# Linear (+ReLU +dropout): region projections
# ------------------------------------------------------------------------------------------------
class LinearActFn(torch.autograd.Function):
    """y = dropout(relu(x W^T + b)) (modeling.py:1003-1018).  K is zero-padded to a multiple of 8 (TMA strides)."""

    @staticmethod
    def forward(ctx, x, w, b, act, p, training, site):
        _require_cuda(x, "linear input")
        N, K = w.shape
        Kp = (K + 7) // 8 * 8
        x2 = _bf16c(x.reshape(-1, K))
        wc = _bf16c(w)
        if Kp != K:
            x2 = torch.nn.functional.pad(x2, (0, Kp - K))
            wc = torch.nn.functional.pad(wc, (0, Kp - K))
        bc = _bf16c(b)
        M = x2.shape[0]
        y = torch.empty(M, N, device=x.device, dtype=BF16)
        seed = next_seed(f"linear:{site}") if (training and p > 0 and act == 1) else None
        drop = _drop(p, seed)
        L.call("vlpk_linear_fwd", M, N, Kp, x2.data_ptr(), Kp, wc.data_ptr(), Kp, L.ptr(bc), y.data_ptr(), N, act, drop, site, L.stream())
        ctx.save_for_backward(x2, wc, y)
        ctx.meta = (act, p if seed is not None else 0.0, K, Kp, x.shape, w.dtype, None if b is None else b.dtype, x.dtype)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, wc, y = ctx.saved_tensors
        act, p, K, Kp, xshape, wdt, bdt, xdt = ctx.meta
        M, N = y.shape
        dyc = _bf16c(dy.reshape(M, N))
        need_dx = ctx.needs_input_grad[0]
        dpre = torch.empty(M, N, device=dyc.device, dtype=BF16) if act == 1 else None
        dx = torch.empty(M, Kp, device=dyc.device, dtype=BF16) if need_dx else None
        dw = torch.zeros(N, Kp, device=dyc.device, dtype=torch.float32)
        db = torch.zeros(N, device=dyc.device, dtype=torch.float32) if bdt is not None else None
        L.call("vlpk_linear_bwd", M, N, Kp, x2.data_ptr(), Kp, wc.data_ptr(), Kp, y.data_ptr(), N, dyc.data_ptr(), N, L.ptr(dpre), L.ptr(dx), Kp,
               dw.data_ptr(), Kp, L.ptr(db), act, float(p), L.stream())
        gx = dx[:, :K].reshape(xshape).to(xdt) if need_dx else None
        return gx, dw[:, :K].to(wdt), (db.to(bdt) if db is not None else None), None, None, None, None


# ------------------------------------------------------------------------------------------------
# Embeddings
# ------------------------------------------------------------------------------------------------
class EmbedFn(torch.autograd.Function):
    """BertEmbeddings.forward (modeling.py:217-241)."""

    @staticmethod
    def forward(ctx, vis, vpe, word_w, pos_w, type_w, ln_g, ln_b, ids, tt, pos, vis_input, R, p, training, dp_hook=None):
        _require_cuda(ids, "input_ids")
        B, Lq = ids.shape
        H = word_w.shape[1]
        visc, vpec = (_bf16c(vis), _bf16c(vpe)) if vis_input else (None, None)
        tabs = [_bf16c(t) for t in (word_w, pos_w, type_w, ln_g, ln_b)]
        ids = ids.contiguous()
        tt = None if tt is None else tt.contiguous()
        pos = None if pos is None else pos.contiguous()
        y = torch.empty(B, Lq, H, device=ids.device, dtype=BF16)
        stats = torch.empty(B * Lq, 2, device=ids.device, dtype=torch.float32)
        seed = next_seed("embed") if (training and p > 0) else None
        drop = _drop(p, seed)
        L.call("vlpk_embed_fwd", B, Lq, H, R, 1 if vis_input else 0, ids.data_ptr(), L.ptr(tt), L.ptr(pos), tabs[0].data_ptr(), tabs[1].data_ptr(),
               tabs[2].data_ptr(), L.ptr(visc), L.ptr(vpec), tabs[3].data_ptr(), tabs[4].data_ptr(), y.data_ptr(), stats.data_ptr(), drop, 1 << 20,
               L.stream())
        ctx.saved = (visc, vpec, tabs, ids, tt, pos, stats)
        ctx.dp_hook = dp_hook
        ctx.meta = (vis_input, R, p if seed is not None else 0.0, seed, [t.dtype for t in (vis, vpe, word_w, pos_w, type_w, ln_g, ln_b)] if vis_input
                    else [None, None] + [t.dtype for t in (word_w, pos_w, type_w, ln_g, ln_b)])
        return y

    @staticmethod
    def backward(ctx, dy):
        visc, vpec, tabs, ids, tt, pos, stats = ctx.saved
        vis_input, R, p, seed, dts = ctx.meta
        B, Lq = ids.shape
        H = tabs[0].shape[1]
        dev = ids.device
        dyc = _bf16c(dy)
        dz = torch.empty(B, Lq, H, device=dev, dtype=BF16)
        dg = torch.zeros(H, device=dev, dtype=torch.float32)
        db = torch.zeros(H, device=dev, dtype=torch.float32)
        drop = _drop(p, seed)
        L.call("vlpk_embed_bwd", B, Lq, H, R, 1 if vis_input else 0, ids.data_ptr(), L.ptr(tt), L.ptr(pos), tabs[0].data_ptr(), tabs[1].data_ptr(),
               tabs[2].data_ptr(), L.ptr(visc), L.ptr(vpec), tabs[3].data_ptr(), stats.data_ptr(), dyc.data_ptr(), dz.data_ptr(), dg.data_ptr(),
               db.data_ptr(), drop, 1 << 20, L.stream())
        d_vis = dz[:, 1:R + 1] if vis_input else None
        # pre-LN gradient -> tables: region rows feed the projections, the B x (L - R) looked-up rows are scattered by
        # vlpk_embed_tables_bwd (csrc/tables.cu: touched-rows-only word/position scatter, segmented token-type sums)
        V, P, T = tabs[0].shape[0], tabs[1].shape[0], tabs[2].shape[0]
        hook = ctx.dp_hook
        if hook is not None and hook.wants_embedding_rows():
            # data parallelism: hand the looked-up rows (23 per sample) to the reducer, which all-gathers them and adds every rank's rows
            # to the word / position gradients itself — instead of all-reducing two dense tables after backward (vlp_b200/dp.py)
            if vis_input:
                keep = torch.cat((torch.zeros(1, dtype=torch.long, device=dev), torch.arange(R + 1, Lq, device=dev)))
                rows = dz[:, keep].reshape(-1, H)
                ids_tab = ids[:, keep].reshape(-1)
                pos_tab = (pos[:, keep] if pos is not None else keep.unsqueeze(0).expand(B, -1)).reshape(-1)
            else:
                rows, ids_tab = dz.reshape(-1, H), ids.reshape(-1)
                pos_tab = (pos if pos is not None else torch.arange(Lq, device=dev).unsqueeze(0).expand(B, -1)).reshape(-1)
            d_word, d_pos = hook.on_embedding_rows(ids_tab.contiguous(), pos_tab.contiguous(), rows.contiguous(), V, P)
            d_type = torch.zeros(T, H, device=dev, dtype=torch.float32)
            L.call("vlpk_embed_tables_bwd", B, Lq, H, R, 1 if vis_input else 0, ids.data_ptr(), L.ptr(tt), L.ptr(pos), dz.data_ptr(), V, P, T,
                   None, None, None, d_type.data_ptr(), L.stream())
            out = [None if d_vis is None else d_vis.to(dts[0]), None if d_vis is None else d_vis.to(dts[1]),
                   None if d_word is None else d_word.to(dts[2]), None if d_pos is None else d_pos.to(dts[3]), d_type.to(dts[4]), dg.to(dts[5]),
                   db.to(dts[6])]
            return tuple(out) + (None,) * 8
        d_word = torch.empty(V, H, device=dev, dtype=BF16)
        scratch = torch.empty(V, H, device=dev, dtype=torch.float32)
        d_pos = torch.zeros(P, H, device=dev, dtype=torch.float32)
        d_type = torch.zeros(T, H, device=dev, dtype=torch.float32)
        L.call("vlpk_embed_tables_bwd", B, Lq, H, R, 1 if vis_input else 0, ids.data_ptr(), L.ptr(tt), L.ptr(pos), dz.data_ptr(), V, P, T,
               d_word.data_ptr(), scratch.data_ptr(), d_pos.data_ptr(), d_type.data_ptr(), L.stream())
        out = [None if d_vis is None else d_vis.to(dts[0]), None if d_vis is None else d_vis.to(dts[1]), d_word.to(dts[2]), d_pos.to(dts[3]),
               d_type.to(dts[4]), dg.to(dts[5]), db.to(dts[6])]
        return tuple(out) + (None,) * 8


# ------------------------------------------------------------------------------------------------
# masked-LM head tail: tied decoder + bias + per-position cross-entropy
# ------------------------------------------------------------------------------------------------
class DecoderCEFn(torch.autograd.Function):
    """loss[r] = CE(h[r] W^T + bias, labels[r]) (modeling.py:478-482 + 1108-1109) without fp32 logits; also returns the bf16
    logits [R,V] (non-differentiable view, kept for `last_prediction_scores`)."""

    @staticmethod
    def forward(ctx, h, w, bias, labels, dp_hook=None):
        _require_cuda(h, "decoder input")
        ctx.dp_hook = dp_hook
        R, H = h.shape
        V = w.shape[0]
        Vp = (V + 7) // 8 * 8
        hc, wc = _bf16c(h), _bf16c(w)
        bias_pad = torch.zeros(Vp, device=h.device, dtype=BF16)
        bias_pad[:V] = bias
        labels = labels.contiguous()
        logits = torch.empty(R, Vp, device=h.device, dtype=BF16)
        lse = torch.empty(R, device=h.device, dtype=torch.float32)
        loss = torch.empty(R, device=h.device, dtype=torch.float32)
        L.call("vlpk_decoder_ce_fwd", R, V, H, hc.data_ptr(), wc.data_ptr(), bias_pad.data_ptr(), labels.data_ptr(), logits.data_ptr(),
               lse.data_ptr(), loss.data_ptr(), L.stream())
        ctx.save_for_backward(hc, wc, labels, logits, lse)
        ctx.meta = (h.dtype, w.dtype, bias.dtype)
        scores = logits[:, :V]
        ctx.mark_non_differentiable(scores)
        return loss, scores

    @staticmethod
    def backward(ctx, dloss, _dscores):
        hc, wc, labels, logits, lse = ctx.saved_tensors
        hdt, wdt, bdt = ctx.meta
        R, H = hc.shape
        V = wc.shape[0]
        Vp = logits.shape[1]
        dev = hc.device
        dl = dloss.to(torch.float32).contiguous()
        dlogits = torch.empty(R, Vp, device=dev, dtype=BF16)
        dh = torch.zeros(R, H, device=dev, dtype=torch.float32)
        dw = torch.empty(V, H, device=dev, dtype=BF16)
        dbias = torch.zeros(Vp, device=dev, dtype=torch.float32)
        L.call("vlpk_decoder_ce_bwd", R, V, H, hc.data_ptr(), wc.data_ptr(), labels.data_ptr(), logits.data_ptr(), lse.data_ptr(), dl.data_ptr(),
               dlogits.data_ptr(), dh.data_ptr(), dw.data_ptr(), dbias.data_ptr(), L.stream())
        dwp = dw.to(wdt)
        if ctx.dp_hook is not None:
            ctx.dp_hook.on_decoder_weight_grad(dwp)      # data parallelism: this (tied) gradient is complete FIRST — reduce it now
        return dh.to(hdt), dwp, dbias[:V].to(bdt), None, None


def layer_cached_fwd(hidden, kv_cache, pos, mask_bits, heads, I, params):
    """BertLayer.forward for decode with a persistent K/V cache (vlpk_layer_cached_fwd): `hidden` [B, Lq, H] are the new rows, `kv_cache`
    [B, rows, 2H] bf16 holds this layer's key | value projections of the `pos` rows already seen; the new rows' K | V are appended at
    [pos, pos + Lq).  Inference only."""
    x = _bf16c(hidden)
    B, Lq, H = x.shape
    if not (kv_cache.dtype == BF16 and kv_cache.is_contiguous() and kv_cache.shape[0] == B and kv_cache.shape[2] == 2 * H):
        raise RuntimeError("vlp_b200: kv_cache must be a contiguous bf16 [B, rows, 2H] tensor")
    pk = [_bf16c(p) for p in params]
    acts = _Acts(1, B, Lq, H, heads, I, x.device, Lkv=Lq)            # acts.kv: scratch for the new rows' K | V
    shape = L.VlpkShape(B, Lq, pos + Lq, H, heads, I)
    ws = _weight_structs(pk, 1)
    L.call("vlpk_layer_cached_fwd", C.byref(shape), ws, x.data_ptr(), kv_cache.data_ptr(), kv_cache.shape[1], int(pos), mask_bits.data_ptr(),
           mask_bits.shape[1], acts.structs, 0, L.stream())
    return acts.y[0]
