"""Beam search for BertForSeq2SeqDecoder (semantics of the reference's modeling.py:1256-1494).

Same algorithm and return format (a `traces` dict of padded tensors: pred_seq, scores, wids, ptrs), with the
beam bookkeeping kept on the device and the back-pointer computed by integer floor division — the reference's
`torch.div(k_ids, K)` (:1317) yields floats on torch >= 1.6 and breaks `gather`.
Every step runs the incremental fused layers (q rows = new token + [MASK], kv rows = cached prefix + q).
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
    curr_ids = input_ids
    mask_ids = input_ids[:, :1] * 0 + dec.mask_word_id
    total_scores, beam_eos, step_ids, step_ptrs = [], [], [], []
    partial, forbid = None, None
    next_pos = in_len
    while next_pos < out_len:
        cl = curr_ids.shape[1]
        st = next_pos - cl
        x_ids = torch.cat((curr_ids, mask_ids), dim=1)
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
        first = prev_emb is None
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
            prev_emb = _expand_beams(new_emb[:, :-1, :], K)
            prev_layers = [_expand_beams(x[:, :-1, :], K) for x in new_layers]
            token_type_ids, position_ids = _expand_beams(token_type_ids, K), _expand_beams(position_ids, K)
            attention_mask, mask_ids = _expand_beams(attention_mask, K), _expand_beams(mask_ids, K)
            vis_feats_k, vis_pe_k = vis_feats, vis_pe                      # regions only enter at step 0
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

    # host-side back-tracking, identical selection rule to the reference (:1431-1472)
    ts = [x.tolist() for x in total_scores]
    si = [x.tolist() for x in step_ids]
    sp = [x.tolist() for x in step_ptrs]
    traces = {"pred_seq": [], "scores": [], "wids": [], "ptrs": []}
    for b in range(B):
        scores = [x[b] for x in ts]
        wids_list = [x[b] for x in si]
        ptrs = [x[b] for x in sp]
        traces["scores"].append(scores)
        traces["wids"].append(wids_list)
        traces["ptrs"].append(ptrs)
        last = len(scores) - 1
        for i, w in enumerate(wids_list):
            if all(x == dec.eos_id for x in w):
                last = i
                break
        best, frame, pos = -math.inf, -1, -1
        for fid in range(last + 1):
            for i, w in enumerate(wids_list[fid]):
                if w == dec.eos_id or fid == last:
                    s = scores[fid][i] + dec.length_penalty * (fid + 1)
                    if s > best:
                        best, frame, pos = s, fid, i
        if frame == -1:
            traces["pred_seq"].append([0])
        else:
            seq = [wids_list[frame][pos]]
            for fid in range(frame, 0, -1):
                pos = ptrs[fid][pos]
                seq.append(wids_list[fid - 1][pos])
            traces["pred_seq"].append(seq[::-1])
    out = {}
    for k, lst in traces.items():
        dt = torch.float if k == "scores" else torch.long
        tens = [torch.tensor(x, dtype=dt) for x in lst]
        padded = tens[0].new_zeros((len(tens), out_len) + tuple(tens[0].shape[1:]))
        for i, t in enumerate(tens):
            padded[i, :t.shape[0]] = t
        out[k] = padded.to(dev)
    return out
