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


def test_training_forward_on_cpu_fails_loudly():
    m = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=1)).train()
    with pytest.raises(N.NativeError, match="no CPU fallback"):
        m(torch.zeros(1, 2, 3, 224, 224))


def _reference_classes():
    """The real reference, from the mount or from the staged copy (baseline/stage_reference.py)."""
    import sys
    sys.path.insert(0, ROOT)
    from baseline import stage_reference as SR
    if not SR.stage(quiet=True):
        pytest.skip("reference neither mounted at /root/reference nor staged under baseline/_ref")
    return SR.import_reference()


@pytest.mark.parametrize("lora", [False, True])
def test_state_dict_names_match_the_real_reference(lora):
    """Direct comparison with the reference's own classes (not via the oracle's name table)."""
    RefConfig, RefModel = _reference_classes()
    kw = dict(num_hidden_layers=2, enable_causal_temporal=True, add_lora_spatial=lora)
    ref = {k: tuple(v.shape) for k, v in RefModel(RefConfig(**kw)).state_dict().items()}
    ours = {k: tuple(v.shape) for k, v in TimesformerMultiTaskingModelSigLIP(StreamformerConfig(**kw)).state_dict().items()}
    assert ours.keys() == ref.keys(), sorted(set(ours) ^ set(ref))
    assert ours == ref


def test_standalone_submodules_expose_reference_names_and_signatures():
    """downstream/AR and the OVIS adapter build the sub-modules on their own (…video_classification.py:42-56):
    same constructor, same parameter names under the same attribute paths, reference forward signatures."""
    import inspect
    from streamformer_b200 import modeling_timesformer_siglip as M
    cfg = StreamformerConfig(num_hidden_layers=2, enable_causal_temporal=True)
    full = {k for k in TimesformerMultiTaskingModelSigLIP(cfg).state_dict()}
    for prefix, mod in [("embeddings.", M.TimesformerEmbeddingsSigLIP(cfg)), ("encoder.", M.TimesformerEncoder(cfg)),
                        ("head.", M.TimesformerSiglipMultiheadAttentionPoolingHead(cfg)),
                        ("encoder.layer.1.", M.TimesformerLayerSigLIP(cfg, 1))]:
        names = {prefix + k for k in mod.state_dict()}
        assert names and names <= full, sorted(names - full)[:5]
        assert mod._sf_prefix == prefix
        assert hasattr(mod, "rebind_weights")
    sig = inspect.signature(M.TimesformerEncoder.forward)
    assert list(sig.parameters)[:6] == ["self", "hidden_states", "num_frames", "output_attentions", "output_hidden_states",
                                        "return_dict"]                       # …siglip.py:1019-1026
    enc = M.TimesformerEncoder(cfg)
    with torch.no_grad(), pytest.raises(N.NativeError, match="no CPU fallback"):
        enc(torch.zeros(1, 196 * 2, 768), num_frames=2)


def test_weight_change_detection_host_logic():
    """_Engine.sync_weights re-binds on version bumps and on mark_dirty (._apply, load_state_dict,
    rebind_weights) and is otherwise a cheap no-op: checked on the host with a stub engine."""
    from streamformer_b200 import modeling_timesformer_siglip as M
    m = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(num_hidden_layers=1, enable_causal_temporal=True))

    class Stub(M._Engine):
        def __init__(self):
            self.prefix, self.params, self.versions, self.names, self.binds, self.pos_key = "", None, None, [], 0, None

        def sync_weights(self, root):   # the change-detection half of the real method
            if self.params is not None and [p._version for p in self.params] == self.versions:
                return
            if self.params is None:
                named = list(root.named_parameters())
                self.names, self.params = [n for n, _ in named], [p for _, p in named]
            self.versions = [p._version for p in self.params]
            self.binds += 1

    eng = Stub()
    m._sf_engines()[("stub",)] = eng
    eng.sync_weights(m); eng.sync_weights(m)
    assert eng.binds == 1
    with torch.no_grad():
        m.post_layernorm.weight.mul_(2.0)             # versioned in-place update
    eng.sync_weights(m)
    assert eng.binds == 2
    m.post_layernorm.weight.data.mul_(2.0)            # NOT versioned ...
    eng.sync_weights(m)
    assert eng.binds == 2
    m.rebind_weights()                                # ... hence the explicit call
    eng.sync_weights(m)
    assert eng.binds == 3
    m.load_state_dict(m.state_dict())                 # post-hook marks dirty
    assert eng.params is None
    eng.sync_weights(m)
    m.float()                                         # ._apply marks dirty
    assert eng.params is None
    import timeit
    eng.sync_weights(m)
    full = TimesformerMultiTaskingModelSigLIP(StreamformerConfig(enable_causal_temporal=True))
    e2 = Stub(); e2.sync_weights(full)
    per_call = timeit.timeit(lambda: e2.sync_weights(full), number=200) / 200
    assert per_call < 100e-6, f"unchanged-weights check costs {per_call * 1e6:.0f} us per forward (281 tensors)"


def test_pixel_format_detection():
    from streamformer_b200.modeling_timesformer_siglip import _pixel_format
    t, dt, H, W = _pixel_format(torch.zeros(1, 2, 3, 32, 48, dtype=torch.uint8), 3)
    assert (dt, H, W) == (N.SF_U8, 32, 48)
    t, dt, H, W = _pixel_format(torch.zeros(1, 2, 32, 48, 3, dtype=torch.uint8), 3)
    assert (dt, H, W) == (N.SF_U8_HWC, 32, 48)
    t, dt, H, W = _pixel_format(torch.zeros(1, 2, 3, 32, 48, dtype=torch.float64), 3)
    assert (dt, H, W) == (N.SF_F32, 32, 48) and t.dtype == torch.float32
    with pytest.raises(ValueError):
        _pixel_format(torch.zeros(2, 3, 32, 48), 3)
