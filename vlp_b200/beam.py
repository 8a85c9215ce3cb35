"""Beam search for BertForSeq2SeqDecoder (semantics of the reference's modeling.py:1256-1494).

Same algorithm and return format (a `traces` dict of padded tensors: pred_seq, scores, wids, ptrs), with ALL beam
bookkeeping on the device — top-k, back pointers (integer floor division: the reference's `torch.div(k_ids, K)`, :1317, yields
floats on torch >= 1.6 and breaks `gather`), the per-layer K/V caches reordered by the back pointers, and the final
best-hypothesis selection + back-tracking (:1431-1472) as vectorised tensor ops: no host synchronisation inside or after the loop
(the optional duplicate-n-gram filter is the one host-side piece, as in the reference).
Every step runs the fused layers on the two new rows (token, [MASK]) against the K/V caches (`dec.use_kv_cache`), or — reference
data flow — against the re-encoded prefix.
"""
import math

import torch
import torch.nn.functional as F


def _expand_beams(x, K):
    """[B, ...] -> [B*K, ...], each item repeated K times consecutively (reference first_expand, :1326-1333)."""
    return x.unsqueeze(1).expand(x.shape[0], K, *x.shape[1:]).reshape(x.shape[0] * K, *x.shape[1:])


def _reorder(x, back_ptrs, B, K):
    """Select, per batch item, the K parent beams named by back_ptrs [B,K] (reference select_beam_items, :1335-1350)."""
    xs = x.view(B, K, *x.shape[1:])
    idx = back_ptrs.view(B, K, *([1] * (x.dim() - 1))).expand(B, K, *x.shape[1:])
    return torch.gather(xs, 1, idx).reshape(x.shape)


def _dup_ngram_candidates(seq, n, ignore):
    """Words that would complete an n-gram already present in seq (reference get_dup_ngram_candidates, :1390-1406)."""
    if len(seq) < n:
        return []
    tail = seq[-(n - 1):]
    if ignore and any(t in ignore for t in tail):
        return []
    out = set()
    for i in range(len(seq) - (n - 1)):
        if seq[i:i + n - 1] == tail and not (ignore and seq[i + n - 1] in ignore):
            out.add(seq[i + n - 1])
    return sorted(out)


def beam_search(dec, vis_feats, vis_pe, input_ids, token_type_ids, position_ids, attention_mask, task_idx=None):
    K = dec.search_beam_size
    B, in_len = input_ids.shape
    out_len = token_type_ids.shape[1]
    dev = input_ids.device
    prev_emb, prev_layers = None, None
    caches = dec.new_kv_caches(B, dev) if getattr(dec, "use_kv_cache", False) else None
    curr_ids = input_ids
    mask_ids = input_ids[:, :1] * 0 + dec.mask_word_id
    total_scores, beam_eos, step_ids, step_ptrs = [], [], [], []
    partial, forbid = None, None
    next_pos = in_len
    while next_pos < out_len:
        cl = curr_ids.shape[1]
        st = next_pos - cl
        x_ids = torch.cat((curr_ids, mask_ids), dim=1)
        if caches is not None:
            new_emb, last, _ = dec.bert(vis_feats, vis_pe, x_ids, token_type_ids[:, st:next_pos + 1], position_ids[:, st:next_pos + 1],
                                        attention_mask[:, st:next_pos + 1, :next_pos + 1], output_all_encoded_layers=False,
                                        len_vis_input=dec.len_vis_input, kv_caches=caches, cache_pos=st)
            new_layers = [last]
        else:
            new_emb, new_layers, _ = dec.bert(vis_feats, vis_pe, x_ids, token_type_ids[:, st:next_pos + 1], position_ids[:, st:next_pos + 1],
                                              attention_mask[:, st:next_pos + 1, :next_pos + 1], prev_embedding=prev_emb,
                                              prev_encoded_layers=prev_layers, output_all_encoded_layers=True, len_vis_input=dec.len_vis_input)
        scores, _ = dec.cls(new_layers[-1][:, -1:, :], None, task_idx=task_idx)
        logp = F.log_softmax(scores.float(), dim=-1)                      # [B or B*K, 1, V]
        if forbid is not None:
            logp = logp + forbid * -10000.0
        if dec.min_len and (next_pos - in_len + 1 <= dec.min_len):
            logp[:, :, dec.eos_id] = -10000.0
        kk_scores, kk_ids = torch.topk(logp, k=K)                          # [*, 1, K]
        first = (next_pos == in_len)
        if first:
            k_ids = kk_ids.reshape(B, K)
            back = torch.zeros(B, K, dtype=torch.long, device=dev)
            k_scores = kk_scores.reshape(B, K)
        else:
            kk_scores = kk_scores + beam_eos[-1].reshape(B * K, 1, 1) * -10000.0 + total_scores[-1].reshape(B * K, 1, 1)
            k_scores, flat = torch.topk(kk_scores.reshape(B, K * K), k=K)
            back = torch.div(flat, K, rounding_mode="floor")
            k_ids = torch.gather(kk_ids.reshape(B, K * K), 1, flat)
        step_ptrs.append(back)
        step_ids.append(k_ids)
        beam_eos.append((k_ids == dec.eos_id).float())
        total_scores.append(k_scores)
        if first:
            if caches is not None:
                caches = [_expand_beams(c, K).contiguous() for c in caches]
            else:
                prev_emb = _expand_beams(new_emb[:, :-1, :], K)
                prev_layers = [_expand_beams(x[:, :-1, :], K) for x in new_layers]
            token_type_ids, position_ids = _expand_beams(token_type_ids, K), _expand_beams(position_ids, K)
            attention_mask, mask_ids = _expand_beams(attention_mask, K), _expand_beams(mask_ids, K)
        elif caches is not None:
            parent = (back + torch.arange(B, device=dev).unsqueeze(1) * K).reshape(-1)      # beam i continues hypothesis parent[i]
            caches = [c.index_select(0, parent) for c in caches]
        else:
            prev_emb = _reorder(torch.cat((prev_emb, new_emb[:, :-1, :]), dim=1), back, B, K)
            prev_layers = [_reorder(torch.cat((a, b[:, :-1, :]), dim=1), back, B, K) for a, b in zip(prev_layers, new_layers)]
        curr_ids = k_ids.reshape(B * K, 1)
        if dec.forbid_duplicate_ngrams:
            wids, ptrs = k_ids.tolist(), back.tolist()
            if first:
                partial = [[wids[b][k]] for b in range(B) for k in range(K)]
            else:
                partial = [partial[ptrs[b][k] + b * K] + [wids[b][k]] for b in range(B) for k in range(K)]
            forbid = None
            if len(partial[0]) >= dec.ngram_size:
                cands = [_dup_ngram_candidates(s, dec.ngram_size, dec.forbid_ignore_set) for s in partial]
                if any(cands):
                    forbid = torch.zeros(B * K, 1, logp.shape[-1], device=dev)
                    for i, c in enumerate(cands):
                        if c:
                            forbid[i, 0, c] = 1.0
        next_pos += 1

    out = {"pred_seq": backtrack(torch.stack(total_scores), torch.stack(step_ids), torch.stack(step_ptrs), dec.eos_id, dec.length_penalty, out_len)}
    T = len(total_scores)
    for k, t in (("scores", torch.stack(total_scores)), ("wids", torch.stack(step_ids)), ("ptrs", torch.stack(step_ptrs))):
        padded = t.new_zeros((B, out_len, K))
        padded[:, :T] = t.permute(1, 0, 2)
        out[k] = padded
    return out


def backtrack(sc, wi, pt, eos_id, length_penalty, out_len):
    """Best-hypothesis selection + back-tracking, same rule as the reference (:1431-1472), vectorised (runs wherever the traces live):
      last[b]   = first frame whose K words are all [EOS] (else the final frame)
      candidate = (word is [EOS], or frame == last[b]) within frames <= last[b]; score + length_penalty * (frame + 1); FIRST maximum wins
    sc [T,B,K] float scores, wi [T,B,K] word ids, pt [T,B,K] back pointers -> pred_seq [B, out_len] (zero padded)."""
    T, B, K = sc.shape
    dev = sc.device
    frames = torch.arange(T, device=dev).view(T, 1)
    all_eos = (wi == eos_id).all(-1)                                      # [T,B]
    last = torch.where(all_eos.any(0), all_eos.float().argmax(0), torch.full((B,), T - 1, device=dev))      # [B]
    cand = (frames <= last.unsqueeze(0)).unsqueeze(-1) & ((wi == eos_id) | (frames == last.unsqueeze(0)).unsqueeze(-1))
    val = torch.where(cand, sc + length_penalty * (frames + 1).unsqueeze(-1).to(sc.dtype), torch.full_like(sc, -math.inf))
    flat = val.permute(1, 0, 2).reshape(B, T * K)                          # (frame, beam) order = the reference's loop order
    best = flat.argmax(-1)                                                # first maximal value
    frame, pos = torch.div(best, K, rounding_mode="floor"), best % K
    found = torch.isfinite(flat.gather(1, best.unsqueeze(1)).squeeze(1))
    pred = torch.zeros(B, out_len, dtype=torch.long, device=dev)
    bidx = torch.arange(B, device=dev)
    for fid in range(T - 1, -1, -1):                                      # walk the back pointers from `frame` down to 0
        active = found & (fid <= frame)
        tok = wi[fid, bidx, pos]
        pred[:, fid] = torch.where(active, tok, pred[:, fid])
        pos = torch.where(active & (fid > 0), pt[fid, bidx, pos], pos)
    return pred
