"""Diagnostic: per-parameter gradient error of the GPU path vs the fp32 oracle (run live), plus the per-row error of the
gradient flowing into the embedding output.  python tests/diag_grads.py [case ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import make_golden as mg
from oracle import vlp_oracle as O
from vlp_b200 import synth
from vlp_b200 import vlp_modules as vm


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def main(names):
    for name in names:
        dims, B, seed, mode, ragged, tasks = mg.CASES[name]
        sd = synth.make_state_dict(dims, 0, tasks)
        batch = synth.make_batch(dims, B, seed=seed, mode=mode, ragged=ragged, tasks=tasks)
        cfg = vm.BertConfig(dims.vocab, hidden_size=dims.hidden, num_hidden_layers=dims.layers, num_attention_heads=dims.heads,
                            intermediate_size=dims.inter, type_vocab_size=dims.type_vocab, max_position_embeddings=dims.max_pos,
                            hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        model = vm.BertForPreTrainingLossMask(cfg, enable_butd=True, len_vis_input=dims.regions, tasks=tasks)
        model.load_state_dict(sd)
        model = model.cuda().bfloat16().eval()
        cap = {}

        def hook(m, i, o):
            o.retain_grad()
            cap["emb"] = o

        model.bert.embeddings.register_forward_hook(hook)
        b = {k: v.cuda() for k, v in batch.items()}
        ans = b["ans_labels"] if tasks == "vqa2" else None
        losses = model(b["img"].bfloat16(), b["vis_pe"].bfloat16(), b["input_ids"], b["segment_ids"], b["input_mask"], b["masked_ids"], ans,
                       b["is_next"], masked_pos=b["masked_pos"], masked_weights=b["masked_weights"], task_idx=b["task_idx"],
                       vis_masked_pos=b["vis_masked_pos"], mask_image_regions=False, drop_worst_ratio=0.0)
        sum(l.sum() for l in losses).backward()
        # oracle with a retained embedding gradient
        for k, v in sd.items():
            if k != "cls.predictions.decoder.weight":
                v.requires_grad_(True)
        ref_losses, aux = O.pretraining_loss(sd, dims, batch, tasks=tasks, return_all=True)
        aux["embedding"].retain_grad()
        sum(l.sum() for l in ref_losses).backward()
        print(f"== {name}: losses {[float(l) for l in losses]} vs {[float(l) for l in ref_losses]}")
        ge, gr = cap["emb"].grad.float().cpu(), aux["embedding"].grad
        print(f"   d(embedding out) rel {rel(ge, gr):.4f}")
        per_row = ((ge - gr).norm(dim=-1) / (gr.norm(dim=-1) + 1e-20))
        mag = gr.norm(dim=-1)
        for bi in range(B):
            worst = torch.argsort(per_row[bi], descending=True)[:6].tolist()
            print(f"   sample {bi}: worst rows {[(r, round(per_row[bi, r].item(), 3), float('%.2e' % mag[bi, r].item())) for r in worst]}")
        rows = []
        for k, p in model.named_parameters():
            if p.grad is None or sd[k].grad is None:
                continue
            rows.append((rel(p.grad, sd[k].grad), k, sd[k].grad.norm().item()))
        rows.sort(reverse=True)
        for r, k, n in rows[:12]:
            print(f"   {r:8.4f}  |ref|={n:.3e}  {k}")


if __name__ == "__main__":
    main(sys.argv[1:] or list(mg.CASES))
