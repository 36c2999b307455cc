"""End-to-end parity of the CUDA encoder (through the HF-style class -> C ABI -> sm_100a kernels)
against (a) the committed outputs of the real reference (tests/golden/*.npz) and (b) the numpy
oracle recomputed on the same seeded inputs.

Tolerance (stated up-front): bf16 implementation vs fp32 reference with identical
weights — last_hidden_state rel-RMS <= 2e-2 and cosine >= 0.9995; pooler_output rel-RMS <= 1.5e-2
and cosine >= 0.9995 (the reference's own bf16 mode scores 1.2e-2 / 6.9e-3); fp16 <= 3e-3.
Structural invariants are exact: causality, LoRA-at-init == no LoRA, batch independence.
"""
import numpy as np
import pytest
import torch

from oracle import streamformer_oracle as O
from tests.oracle_utils import case_config, case_inputs, cosine, golden_names, load_golden, rel_rms, sub

pytestmark = pytest.mark.gpu

TOL = {
    torch.bfloat16: dict(lhs=2e-2, pool=1.5e-2, cos=0.9995),
    torch.float16: dict(lhs=3e-3, pool=3e-3, cos=0.99999),
}


def build_model(cfg: O.OracleConfig, weights, dtype=torch.bfloat16, **extra):
    from streamformer_b200.modeling_timesformer_siglip import StreamformerConfig, TimesformerMultiTaskingModelSigLIP
    hc = StreamformerConfig(num_hidden_layers=cfg.num_hidden_layers, enable_causal_temporal=cfg.enable_causal_temporal,
                            add_lora_spatial=cfg.add_lora_spatial, num_frames=cfg.num_frames, **extra)
    model = TimesformerMultiTaskingModelSigLIP(hc)
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in weights.items() if k in model.state_dict()}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.endswith(".mask") for m in missing), (missing, unexpected)
    return model.to("cuda", dtype).eval()


def check(name, got, want, rms_tol, cos_tol):
    got = got.float().cpu().numpy() if isinstance(got, torch.Tensor) else got
    assert np.isfinite(got).all(), f"{name}: non-finite values"
    r, c = rel_rms(got, want), cosine(got, want)
    assert r <= rms_tol, f"{name}: rel-RMS {r:.4g} > {rms_tol}"
    assert c >= cos_tol, f"{name}: cosine {c:.6f} < {cos_tol}"
    return r


ROOT_CASES = [n for n in golden_names() if n != "twin_stream"]


@pytest.mark.parametrize("name", ROOT_CASES)
def test_matches_reference_golden_bf16(name):
    case, z = load_golden(name)
    cfg, w, px = case_inputs(case)
    model = build_model(cfg, w)
    with torch.no_grad():
        out = model(torch.from_numpy(px).cuda(), output_hidden_states=True, output_attentions=True)
    t = TOL[torch.bfloat16]
    assert out.last_hidden_state.shape == (case["B"], case["T"], 196, 768)
    assert out.pooler_output.shape == (case["B"], case["T"], 768)
    assert out.last_hidden_state.dtype == torch.bfloat16
    check("pooler_output", out.pooler_output, z["pooler_output"], t["pool"], t["cos"])
    check("last_hidden_state", sub(out.last_hidden_state.float().cpu().numpy()), z["last_hidden_state_sub"], t["lhs"], t["cos"])
    assert len(out.hidden_states) == case["layers"] + 1 and len(out.attentions) == case["layers"]
    check("embedding", out.hidden_states[0].float().cpu().numpy()[:, ::97, ::8], z["embedding_sub"], 1e-2, 0.9999)
    check("hidden_state_1", out.hidden_states[1].float().cpu().numpy()[:, ::97, ::8], z["hidden_state_1_sub"], 1.5e-2, 0.9995)
    att = out.attentions[0].float().cpu().numpy()
    assert att.shape == (case["B"] * case["T"], 12, 196, 196)
    check("attention_0", att[::3, ::5, ::13, :], z["attention_0_sub"], 5e-2, 0.998)


def test_matches_reference_golden_fp16():
    case, z = load_golden("stress_l2_lora")
    cfg, w, px = case_inputs(case)
    model = build_model(cfg, w, torch.float16)
    with torch.no_grad():
        out = model(torch.from_numpy(px).cuda().half())
    t = TOL[torch.float16]
    check("pooler_output fp16", out.pooler_output, z["pooler_output"], t["pool"], t["cos"])
    check("last_hidden_state fp16", sub(out.last_hidden_state.float().cpu().numpy()), z["last_hidden_state_sub"], t["lhs"], t["cos"])


@pytest.mark.parametrize("fold", [True, False])
def test_matches_oracle_full_tensors(fold):
    """Fresh oracle run on the box (no fixture): full tensors, both projection-folding modes."""
    cfg = O.OracleConfig(num_hidden_layers=2)
    w = O.make_weights(cfg, seed=21, style="stress")
    px = O.make_pixels(2, 5, cfg, seed=21)
    ref = O.forward(w, cfg, px)
    model = build_model(cfg, w, fold_temporal_proj=fold)
    with torch.no_grad():
        out = model(torch.from_numpy(px).cuda())
    t = TOL[torch.bfloat16]
    check("pooler_output", out.pooler_output, ref["pooler_output"], t["pool"], t["cos"])
    check("last_hidden_state", out.last_hidden_state, ref["last_hidden_state"], t["lhs"], t["cos"])


@pytest.mark.parametrize("name", ["stress_l2_lora", "long_t128_l1"])
def test_gemm_chain_schedule_matches_golden(name):
    """The opt-in chained schedule (several dependent GEMMs as one persistent launch with in-kernel row
    dependencies) must give the same answer: M = 6272 exercises the short-chain case where a tile's
    producer sits exactly one round earlier on the same worker, M = 25088 the bench geometry."""
    from streamformer_b200 import _native as N
    case, z = load_golden(name)
    cfg, w, px = case_inputs(case)
    model = build_model(cfg, w)
    pxc = torch.from_numpy(px).cuda()
    with torch.no_grad():
        base = model(pxc)
        n0 = N.launch_count()
        model(pxc)
        per_gemm = N.launch_count() - n0
        N.set_option("gemm_chain", 1)
        try:
            n0 = N.launch_count()
            out = model(pxc)
            torch.cuda.synchronize()
            chained = N.launch_count() - n0
        finally:
            N.set_option("gemm_chain", -1)
    assert chained < per_gemm, "the chained schedule did not engage"
    t = TOL[torch.bfloat16]
    check(name + " chain last_hidden_state", sub(out.last_hidden_state.float().cpu().numpy()), z["last_hidden_state_sub"], t["lhs"], t["cos"])
    check(name + " chain pooler_output", out.pooler_output, z["pooler_output"], t["pool"], t["cos"])
    # and against the default schedule (different epilogue warp counts / partial-sum order: tolerance, not bitwise)
    check(name + " chain vs default", out.last_hidden_state, base.last_hidden_state.float().cpu().numpy(), t["lhs"], t["cos"])


def test_full_size_bench_config_properties():
    """BASELINE.json configs[1] at full size (B=8, T=16, 224x224, 12 layers, bf16 — the bench workload):
    the reference's own full-depth golden clip placed inside the batch must come out within tolerance,
    and the size-independent invariants must hold exactly: a clip's result does not depend on its
    batch slot or on the other clips of the launch (different tile / CTA assignment), perturbing the
    last frame leaves every earlier frame bit-identical, and the run is deterministic."""
    case, z = load_golden("full_l12")
    cfg, w, px = case_inputs(case)
    model = build_model(cfg, w)
    g = torch.Generator().manual_seed(99)
    batch = torch.randn(8, 16, 3, 224, 224, generator=g)
    batch[3] = torch.from_numpy(px)[0]
    batch[6] = torch.from_numpy(px)[0]
    batch = batch.cuda()
    with torch.no_grad():
        out = model(batch)
        again = model(batch)
        single = model(batch[3:4])
        pert = batch.clone()
        pert[:, 15] += 0.5
        out_p = model(pert)
    t = TOL[torch.bfloat16]
    check("golden clip inside the batch: pooler_output", out.pooler_output[3:4], z["pooler_output"], t["pool"], t["cos"])
    check("golden clip inside the batch: last_hidden_state", sub(out.last_hidden_state[3:4].float().cpu().numpy()),
          z["last_hidden_state_sub"], t["lhs"], t["cos"])
    assert torch.equal(out.last_hidden_state, again.last_hidden_state) and torch.equal(out.pooler_output, again.pooler_output)
    assert torch.equal(out.last_hidden_state[3], out.last_hidden_state[6]), "same clip, different batch slot"
    # alone (M = 3136) the GEMMs pick other tile shapes, so the LayerNorm statistics are summed in a
    # different order: equal within the dtype tolerance, not bitwise
    check("clip alone vs inside the batch", single.last_hidden_state[0], out.last_hidden_state[3].float().cpu().numpy(), t["lhs"], t["cos"])
    check("clip alone vs inside the batch (pooled)", single.pooler_output[0], out.pooler_output[3].float().cpu().numpy(), t["pool"], t["cos"])
    assert torch.equal(out.last_hidden_state[:, :15], out_p.last_hidden_state[:, :15]), "causality at full size"
    assert torch.equal(out.pooler_output[:, :15], out_p.pooler_output[:, :15])
    assert not torch.equal(out.last_hidden_state[:, 15], out_p.last_hidden_state[:, 15])


def test_fp32_parameters_compute_in_bf16_and_return_fp32():
    cfg = O.OracleConfig(num_hidden_layers=1)
    w = O.make_weights(cfg, seed=22, style="stress")
    px = O.make_pixels(1, 3, cfg, seed=22)
    ref = O.forward(w, cfg, px)
    model = build_model(cfg, w, torch.float32)
    with torch.no_grad():
        out = model(torch.from_numpy(px).cuda())
    assert out.pooler_output.dtype == torch.float32
    check("pooler_output", out.pooler_output, ref["pooler_output"], 1.5e-2, 0.9995)


def test_causality_is_exact():
    cfg = O.OracleConfig(num_hidden_layers=2)
    w = O.make_weights(cfg, seed=23, style="stress")
    model = build_model(cfg, w)
    px = torch.from_numpy(O.make_pixels(2, 16, cfg, seed=23)).cuda()
    with torch.no_grad():
        a = model(px)
        px2 = px.clone()
        px2[:, 15] += 1.0
        b = model(px2)
    assert torch.equal(a.last_hidden_state[:, :15], b.last_hidden_state[:, :15])
    assert torch.equal(a.pooler_output[:, :15], b.pooler_output[:, :15])
    assert not torch.equal(a.last_hidden_state[:, 15], b.last_hidden_state[:, 15])


def test_batch_independence_is_exact():
    cfg = O.OracleConfig(num_hidden_layers=2)
    w = O.make_weights(cfg, seed=24, style="stress")
    model = build_model(cfg, w)
    px = torch.from_numpy(O.make_pixels(3, 8, cfg, seed=24)).cuda()
    with torch.no_grad():
        full = model(px)
        one = model(px[1:2])
    assert torch.equal(full.last_hidden_state[1:2], one.last_hidden_state)
    assert torch.equal(full.pooler_output[1:2], one.pooler_output)


def test_lora_at_init_equals_no_lora():
    cfg = O.OracleConfig(num_hidden_layers=2)
    w = O.make_weights(cfg, seed=25, style="stress")
    base = build_model(cfg, w)
    px = torch.from_numpy(O.make_pixels(1, 4, cfg, seed=25)).cuda()
    with torch.no_grad():
        a = base(px)
        base.add_lora_spatial()           # B = 0 at init (…siglip.py:646-647)
        assert any("lora" in k for k in base.state_dict())
        b = base(px)
    assert torch.equal(a.last_hidden_state, b.last_hidden_state)
    assert torch.equal(a.pooler_output, b.pooler_output)


def test_weights_rebind_after_inplace_update():
    cfg = O.OracleConfig(num_hidden_layers=1)
    w = O.make_weights(cfg, seed=26, style="stress")
    model = build_model(cfg, w)
    px = torch.from_numpy(O.make_pixels(1, 2, cfg, seed=26)).cuda()
    with torch.no_grad():
        a = model(px).pooler_output.clone()
        model.post_layernorm.weight.mul_(1.5)
        b = model(px).pooler_output
    assert not torch.equal(a, b)


@pytest.mark.parametrize("chunks", [[8, 8], [1] * 16, [5, 3, 8]])
def test_streaming_matches_twin_golden_and_full_forward(chunks):
    case, z = load_golden("twin_stream")
    cfg, w, px = case_inputs(case)
    model = build_model(cfg, w)
    pxc = torch.from_numpy(px).cuda()
    with torch.no_grad():
        full = model(pxc)
        cache = model.new_kv_cache(batch_size=case["B"], max_frames=16)
        parts, pools, pos = [], [], 0
        for n in chunks:
            r = model(pxc[:, pos:pos + n], past_key_values=cache, use_cache=True)
            assert r.past_key_values is cache and cache.get_seq_length() == pos + n
            parts.append(r.last_hidden_state)
            pools.append(r.pooler_output)
            pos += n
    got = torch.cat(parts, dim=1)
    t = TOL[torch.bfloat16]
    # vs the reference twin's own streamed output, and vs the reference's one-shot output
    check("stream vs twin golden", sub(got.float().cpu().numpy()), z["stream_0_last_hidden_state_sub"], t["lhs"], t["cos"])
    check("stream vs one-shot golden", sub(got.float().cpu().numpy()), z["last_hidden_state_sub"], t["lhs"], t["cos"])
    # streaming == full forward within 2x the dtype tolerance
    check("stream vs own full", got, full.last_hidden_state.float().cpu().numpy(), 2 * t["lhs"], t["cos"])
    check("stream pooler vs own full", torch.cat(pools, dim=1), full.pooler_output.float().cpu().numpy(), 2 * t["pool"], t["cos"])


def test_streaming_graph_replay_is_exact_and_active():
    """From the second step on a streaming step is replayed from one captured CUDA graph whose kernels
    read the stream position from a device counter: the replays must equal (bitwise) what direct
    launches produce, a second stream after reset() must reproduce the first, and an interleaved
    non-streaming forward (different workspace use) must not disturb the stream."""
    cfg = O.OracleConfig(num_hidden_layers=2, num_frames=24)
    w = O.make_weights(cfg, seed=31, style="stress")
    model = build_model(cfg, w)
    Tt = 24
    px = torch.from_numpy(O.make_pixels(2, Tt, cfg, seed=31)).cuda()
    with torch.no_grad():
        # direct launches: output_hidden_states=True bypasses the graph path
        ref_cache = model.new_kv_cache(batch_size=2, max_frames=Tt)
        ref = [model(px[:, i:i + 1], past_key_values=ref_cache, output_hidden_states=True) for i in range(Tt)]
        assert ref_cache.graph_launches == 0
        cache = model.new_kv_cache(batch_size=2, max_frames=Tt)
        got = []
        for i in range(Tt):
            got.append(model(px[:, i:i + 1], past_key_values=cache))
            if i == 10:
                model(px[:, :4])          # a one-shot forward in between reuses the same workspace
        assert cache.graph_launches >= Tt - 3, "streaming steps were not served by the CUDA graph"
        for i in range(Tt):
            assert torch.equal(got[i].last_hidden_state, ref[i].last_hidden_state), f"step {i}"
            assert torch.equal(got[i].pooler_output, ref[i].pooler_output), f"step {i} pooler"
        cache.reset()
        again = [model(px[:, i:i + 1], past_key_values=cache).last_hidden_state for i in range(Tt)]
        for i in range(Tt):
            assert torch.equal(again[i], ref[i].last_hidden_state), f"second stream, step {i}"


def test_streaming_cache_overflow_and_reset():
    from streamformer_b200 import _native as N
    cfg = O.OracleConfig(num_hidden_layers=1)
    w = O.make_weights(cfg, seed=27)
    model = build_model(cfg, w)
    px = torch.from_numpy(O.make_pixels(1, 3, cfg, seed=27)).cuda()
    cache = model.new_kv_cache(batch_size=1, max_frames=4)
    with torch.no_grad():
        model(px, past_key_values=cache)
        with pytest.raises(N.NativeError, match="overflow"):
            model(px, past_key_values=cache)
        cache.reset()
        assert cache.get_seq_length() == 0
        model(px, past_key_values=cache)


def test_streaming_beyond_num_frames_fixed_horizon():
    """64 appended frames at B=1 with num_frames=16: with time_horizon=T_total streaming equals the
    one-shot T_total forward (nearest time-embedding map over a fixed horizon)."""
    cfg = O.OracleConfig(num_hidden_layers=1)
    w = O.make_weights(cfg, seed=28, style="stress")
    model = build_model(cfg, w)
    Tt = 40
    px = torch.from_numpy(O.make_pixels(1, Tt, cfg, seed=28)).cuda()
    with torch.no_grad():
        full = model(px)
        cache = model.new_kv_cache(batch_size=1, max_frames=Tt, time_horizon=Tt)
        parts = [model(px[:, i:i + 1], past_key_values=cache).last_hidden_state for i in range(Tt)]
    got = torch.cat(parts, dim=1)
    check("long stream vs full", got, full.last_hidden_state.float().cpu().numpy(), 4e-2, 0.9995)


def test_block_level_api_matches_forward():
    """embeddings(...) -> encoder.layer[i](x, T)[0] -> post_layernorm/head: the AR / OVIS call pattern."""
    cfg = O.OracleConfig(num_hidden_layers=2)
    w = O.make_weights(cfg, seed=29, style="stress")
    model = build_model(cfg, w)
    px = torch.from_numpy(O.make_pixels(2, 4, cfg, seed=29)).cuda()
    with torch.no_grad():
        full = model(px, output_hidden_states=True)
        x, gh, gw = model.embeddings(px, return_size=True)
        assert (gh, gw) == (14, 14)
        assert torch.equal(x, full.hidden_states[0])
        for i, blk in enumerate(model.encoder.layer):
            x = blk(x, 4, output_attentions=False)[0]
            # same kernels; only the LayerNorm row statistics of the block's input are summed in a
            # different order (one rowstats pass here, GEMM-epilogue partials inside forward), so
            # some outputs round to the neighbouring bf16 value (and layer 2 starts from those)
            ref = full.hidden_states[i + 1].float()
            diff = (x.float() - ref).abs()
            rel = float(diff.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
            assert float(diff.max()) <= 2.0 ** -5 * float(ref.abs().max()), float(diff.max())
            assert rel <= 2.0 ** -8, rel
        pooled = model.head(full.last_hidden_state.reshape(8, 196, 768))
    assert torch.equal(pooled.reshape(2, 4, 768), full.pooler_output)


def test_variable_resolution_runs():
    """Non-square input -> bicubic position table (…siglip.py:380-411)."""
    cfg = O.OracleConfig(num_hidden_layers=1)
    w = O.make_weights(cfg, seed=30)
    model = build_model(cfg, w)
    px = torch.randn(1, 2, 3, 224, 448, device="cuda")
    with torch.no_grad():
        out = model(px)
    assert out.last_hidden_state.shape == (1, 2, 392, 768)
    assert torch.isfinite(out.last_hidden_state.float()).all()
