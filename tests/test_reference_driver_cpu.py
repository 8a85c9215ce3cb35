"""Boundary proof: the reference's OWN training driver, `vlp/run_img2txt_dist.py`, executed UNMODIFIED — argument
parsing, tokenizer, `Preprocess4Seq2seq` / `Img2txtDataset` loader, `BertForPreTrainingLossMask.from_pretrained(...)`, the training-
loop body (:462-586: `model(conv_feats, vis_pe, input_ids, ...)`, `loss.backward()`, `BertAdam.step()`), checkpoint save — with
`vlp_b200.install(optimizer=True)` serving `pytorch_pretrained_bert.modeling` / `.optimization`.

What is stubbed is only what is OUTSIDE the hot path and absent from this image: `h5py` (feature files -> synthetic arrays),
`pycocoevalcap` (SCST reward, unused), `boto3` (download helper).  There is no GPU in the build container, so the library calls are
replaced by their ctypes prototype check (tools/abi_cases.dry_run): every call the driver triggers is marshalled against
include/vlpk.h, values are meaningless.  Numerics of the same module surface are covered by the GPU parity tests.

Skipped where /root/reference does not exist (the GPU box)."""
import json
import os
import runpy
import sys
import types

import numpy as np
import pytest
import torch

REF = os.environ.get("VLP_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "vlp")), reason="reference checkout not present")

VOCAB = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + [f"w{i}" for i in range(195)]


class _H5File:
    """h5py.File stand-in: any key -> synthetic Detectron outputs of the shape the loader expects (seq2seq_loader.py:322-333)."""

    def __init__(self, path, mode="r"):
        self.path = path

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __getitem__(self, key):
        rng = np.random.RandomState(abs(hash((self.path, key))) % (2 ** 31))
        import re
        base = os.path.basename(self.path)
        if re.search(r"_feat\d{3}\.h5$", base):
            return np.maximum(rng.randn(100, 2048), 0).astype(np.float32)
        if re.search(r"_cls\d{3}\.h5$", base):
            return rng.rand(100, 1601).astype(np.float32)
        box = rng.rand(100, 6).astype(np.float32)
        box[:, 2:4] += box[:, 0:2]
        return box


def _write_inputs(tmp):
    model_dir = os.path.join(tmp, "bert-tiny")
    os.makedirs(model_dir)
    json.dump({"vocab_size": len(VOCAB), "hidden_size": 128, "num_hidden_layers": 2, "num_attention_heads": 2, "intermediate_size": 512,
               "hidden_act": "gelu", "hidden_dropout_prob": 0.1, "attention_probs_dropout_prob": 0.1, "max_position_embeddings": 512,
               "type_vocab_size": 2, "initializer_range": 0.02}, open(os.path.join(model_dir, "bert_config.json"), "w"))
    open(os.path.join(model_dir, "vocab.txt"), "w").write("\n".join(VOCAB) + "\n")
    images = [{"split": "train", "filename": f"COCO_train2014_{i:012d}.jpg", "filepath": "train2014",
               "sentences": [{"raw": " ".join(f"w{(7 * i + j) % 190}" for j in range(5 + i))}]} for i in range(4)]
    src = os.path.join(tmp, "dataset_coco.json")
    json.dump({"images": images}, open(src, "w"))
    return model_dir, src


def test_reference_training_driver_runs_unmodified_through_install(tmp_path, monkeypatch):
    from tools import abi_cases
    from vlp_b200 import install as vinstall
    from oracle import ref_shim

    tmp = str(tmp_path)
    model_dir, src = _write_inputs(tmp)
    saved_modules = dict(sys.modules)
    saved_path = list(sys.path)
    try:
        # third-party modules the image lacks, none of them on the hot path
        h5 = types.ModuleType("h5py")
        h5.File = _H5File
        sys.modules["h5py"] = h5
        for name in ("pycocoevalcap", "pycocoevalcap.cider", "pycocoevalcap.cider.cider"):
            sys.modules[name] = types.ModuleType(name)
        sys.modules["pycocoevalcap.cider.cider"].Cider = type("Cider", (), {"__init__": lambda self, *a, **k: None})
        ref_shim.import_reference_modeling()                 # registers the package object (tokenization stays the reference's) + boto3 stubs
        for name in ("pytorch_pretrained_bert.modeling", "pytorch_pretrained_bert.optimization"):
            sys.modules.pop(name, None)
        served = vinstall.install(optimizer=True)
        assert sys.modules["pytorch_pretrained_bert.modeling"] is served
        sys.path.insert(0, REF)
        out_dir = os.path.join(tmp, "out")
        argv = ["run_img2txt_dist.py", "--do_train", "--enable_butd", "--from_scratch", "--new_segment_ids", "--bert_model", model_dir,
                "--output_dir", out_dir, "--src_file", src, "--image_root", tmp, "--file_valid_jpgs", os.path.join(tmp, "valid.json"),
                "--dataset", "coco", "--split", "train", "--train_batch_size", "2", "--num_train_epochs", "1", "--num_workers", "0",
                "--len_vis_input", "100", "--max_len_b", "20", "--max_pred", "3", "--mask_prob", "0.7", "--learning_rate", "3e-5",
                "--no_cuda"]
        monkeypatch.setattr(sys, "argv", argv)
        with abi_cases.dry_run() as calls:
            runpy.run_path(os.path.join(REF, "vlp", "run_img2txt_dist.py"), run_name="__main__")
    finally:
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_modules:
                del sys.modules[k]
        sys.modules.update(saved_modules)

    # 4 captions / batch 2 = 2 optimisation steps, each: 3 region projections, embeddings, mask pack, fused encoder fwd, fused head,
    # their backward, and ONE fused optimizer launch for all tensors
    steps = calls.count("vlpk_bertadam_step")
    assert steps == 2
    assert calls.count("vlpk_encoder_fwd") == steps and calls.count("vlpk_encoder_bwd") == steps
    assert calls.count("vlpk_linear_fwd") == 3 * steps and calls.count("vlpk_linear_bwd") == 3 * steps
    assert calls.count("vlpk_decoder_ce_fwd") == steps and calls.count("vlpk_embed_tables_bwd") == steps
    # the driver's checkpoint: same parameter names (and shapes) as the REFERENCE class built from the same config
    ckpt = torch.load(os.path.join(out_dir, "model.1.bin"))
    from vlp_b200 import synth
    dims = synth.VlpDims(vocab=len(VOCAB), hidden=128, layers=2, heads=2, inter=512, type_vocab=6)
    ref_model = ref_shim.build_reference_model(dims, {k: v for k, v in ckpt.items()})      # raises on any missing / unexpected key
    ref_sd = ref_model.state_dict()
    assert set(ref_sd) == set(ckpt)
    assert all(tuple(ref_sd[k].shape) == tuple(ckpt[k].shape) for k in ckpt)
    assert os.path.exists(os.path.join(out_dir, "opt.json"))
