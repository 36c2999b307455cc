"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every declared
symbol, the HF-style class exposes the reference's state-dict names, save/from_pretrained round-trips,
and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import streamformer_oracle as O
from streamformer_b200 import _native as N
from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = N.load()
    header = open(os.path.join(ROOT, "include", "streamformer_b200.h")).read()
    declared = set(re.findall(r"\b(sf_[a-z0-9_]+)\s*\(", header))
    declared -= {"sf_config", "sf_status"}
    assert declared, "no declarations parsed"
    assert declared == set(N.EXPORTED_SYMBOLS), declared ^ set(N.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by {N.LIB_PATH}"
    assert b"sm_100a" in lib.sf_version()


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors must have the sizes gcc computes from include/streamformer_b200.h."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "streamformer_b200.h"\nint main(void){printf("%zu %zu %zu\\n",'
                   'sizeof(sf_config),sizeof(sf_weight_desc),sizeof(sf_gemm_epilogue));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(N.SfConfig), ctypes.sizeof(N.SfWeightDesc), ctypes.sizeof(N.SfGemmEpilogue)]


def test_state_dict_names_match_reference():
    for lora in (False, True):
        cfg = StreamformerConfig(num_hidden_layers=2, enable_causal_temporal=True, add_lora_spatial=lora)
        model = TimesformerMultiTaskingModelSigLIP(cfg)
        ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        ref = {k: tuple(np.shape(v)) for k, v in
               O.make_weights(O.OracleConfig(num_hidden_layers=2, add_lora_spatial=lora)).items()}
        assert ours.keys() == ref.keys(), sorted(set(ours) ^ set(ref))
        assert ours == ref
    # 12 layers: 281 tensors (+48 with LoRA), SURVEY §8b
    full = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(enable_causal_temporal=True))
    assert len(full.state_dict()) == 281
    assert sum(p.numel() for p in full.parameters()) == 128_350_476


def test_reference_init_values():
    m = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=1, enable_causal_temporal=True))
    assert float(m.encoder.layer[0].temporal_attention_gating) == 0.0         # …siglip.py:896
    assert float(m.embeddings.time_embeddings.abs().max()) == 0.0              # …siglip.py:377
    assert m.base_model_prefix == "timesformer" and m.main_input_name == "pixel_values"
    assert m.encoder.layer[0].attention_type == "divided_space_time"


def test_save_and_from_pretrained_roundtrip(tmp_path):
    cfg = StreamformerConfig(num_hidden_layers=1, enable_causal_temporal=True, add_lora_spatial=True)
    m = TimesformerMultiTaskingModelSigLIP(cfg)
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0, 0.02)
    m.save_pretrained(tmp_path)
    m2 = TimesformerMultiTaskingModelSigLIP.from_pretrained(tmp_path, ignore_mismatched_sizes=True)
    assert m2.config.add_lora_spatial and m2.config.enable_causal_temporal
    a, b = m.state_dict(), m2.state_dict()
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_cpu_forward_fails_loudly():
    m = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=1)).eval()
    with torch.no_grad(), pytest.raises(N.NativeError, match="no CPU fallback"):
        m(torch.zeros(1, 2, 3, 224, 224))


def test_training_forward_is_refused():
    m = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=1)).train()
    with pytest.raises(NotImplementedError, match="forward pass only"):
        m(torch.zeros(1, 2, 3, 224, 224))
