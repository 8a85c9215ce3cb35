"""Host-side checks that run without a GPU: module surface, state_dict contract, C-ABI exports, error behaviour."""
import ctypes
import os
import re

import pytest
import torch

from vlp_b200 import _lib, synth
from vlp_b200 import vlp_modules as vm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_config(d):
    return vm.BertConfig(d.vocab, hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads, intermediate_size=d.inter,
                         type_vocab_size=d.type_vocab, max_position_embeddings=d.max_pos)


@pytest.mark.parametrize("tasks", ["img2txt", "vqa2"])
def test_state_dict_keys_match_reference_contract(tasks):
    """Parameter names/shapes are the checkpoint contract; synth.state_dict_keys lists the reference's."""
    d = synth.TINY
    model = vm.BertForPreTrainingLossMask(make_config(d), enable_butd=True, len_vis_input=d.regions, tasks=tasks)
    sd = model.state_dict()
    want = {k: tuple(s) for k, s, _ in synth.state_dict_keys(d, tasks)}
    want["cls.predictions.decoder.weight"] = want["bert.embeddings.word_embeddings.weight"]
    assert set(sd.keys()) == set(want.keys())
    for k, shape in want.items():
        assert tuple(sd[k].shape) == shape, k
    res = model.load_state_dict(synth.make_state_dict(d, 0, tasks), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert model.cls.predictions.decoder.weight is model.bert.embeddings.word_embeddings.weight


def test_decoder_surface():
    d = synth.TINY
    m = vm.BertForSeq2SeqDecoder(make_config(d), mask_word_id=103, eos_id=102, enable_butd=True, len_vis_input=d.regions)
    keys = set(m.state_dict().keys())
    assert "vis_embed.0.weight" in keys and "bert.encoder.layer.1.output.LayerNorm.bias" in keys and "cls.predictions.bias" in keys


def test_unsupported_configs_raise():
    with pytest.raises(NotImplementedError):
        vm.BertLayer(vm.BertConfig(100, hidden_size=128, num_attention_heads=2, intermediate_size=512, hidden_act="relu"))
    with pytest.raises(NotImplementedError):
        vm.BertLayer(vm.BertConfig(100, hidden_size=96, num_attention_heads=3, intermediate_size=512))


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of computing somewhere else."""
    d = synth.TINY
    model = vm.BertForPreTrainingLossMask(make_config(d), enable_butd=True, len_vis_input=d.regions)
    b = synth.make_batch(d, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(b["img"], b["vis_pe"], b["input_ids"], b["segment_ids"], b["input_mask"], b["masked_ids"], None, b["is_next"],
              masked_pos=b["masked_pos"], masked_weights=b["masked_weights"], task_idx=b["task_idx"], drop_worst_ratio=0)


def test_library_exports_every_declared_symbol():
    """Every function declared in include/vlpk.h is exported by libvlpk.so and bound in _lib (no compute calls here)."""
    hdr = open(os.path.join(ROOT, "include", "vlpk.h")).read()
    declared = set(re.findall(r"\b(vlpk_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in vlpk.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib().vlpk_version() == 100


def test_install_shadows_reference_import_path():
    import sys
    from vlp_b200 import install
    saved = {k: sys.modules.get(k) for k in ("pytorch_pretrained_bert", "pytorch_pretrained_bert.modeling")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        install.install()
        from pytorch_pretrained_bert.modeling import BertForPreTrainingLossMask, BertForSeq2SeqDecoder  # noqa: F401
        assert BertForPreTrainingLossMask is vm.BertForPreTrainingLossMask
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_flop_model_matches_baseline_md():
    f = synth.flops_per_sample()
    assert abs(f["fwd"] / 1e9 - 22.989) < 0.01 and abs(f["total"] / 1e9 - 67.881) < 0.02
