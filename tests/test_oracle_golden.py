"""Pin the numpy oracle to the REAL reference: every committed golden fixture (outputs of
/root/reference's own forward on seeded weights/inputs, tests/golden/make_golden.py) must be
reproduced by oracle/streamformer_oracle.py in fp32.  Runs on CPU.

Tolerance: both sides are fp32 with different summation orders (torch/oneDNN vs numpy/OpenBLAS):
relative RMS <= 2e-5 per tensor, max abs error <= 2e-4 x the tensor's max magnitude.
"""
import numpy as np
import pytest

from oracle import streamformer_oracle as O
from tests.oracle_utils import case_inputs, golden_names, load_golden, rel_rms, sub

FAST = [n for n in golden_names() if n not in ("full_l12", "long_t128_l1")]
SLOW = [n for n in golden_names() if n in ("full_l12", "long_t128_l1")]


def _check(name, got, want):
    want = np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    r = rel_rms(got, want)
    m = float(np.abs(got - want).max())
    assert r <= 2e-5, f"{name}: rel rms {r:.3g}"
    assert m <= 2e-4 * max(float(np.abs(want).max()), 1.0), f"{name}: max abs {m:.3g}"


def _run(name):
    case, z = load_golden(name)
    cfg, w, px = case_inputs(case)
    if case["model"] == "root":
        r = O.forward(w, cfg, px, output_hidden_states=True, output_attentions=True)
        _check("pooler_output", r["pooler_output"], z["pooler_output"])
        _check("last_hidden_state", sub(r["last_hidden_state"]), z["last_hidden_state_sub"])
        _check("embedding", r["hidden_states"][0][:, ::97, ::8], z["embedding_sub"])
        _check("hidden_state_1", r["hidden_states"][1][:, ::97, ::8], z["hidden_state_1_sub"])
        _check("attention_0", r["attentions"][0][::3, ::5, ::13, :], z["attention_0_sub"])
    else:
        one = O.forward(w, cfg, px)
        _check("twin one-shot", sub(one["last_hidden_state"]), z["last_hidden_state_sub"])
        for ci, chunks in enumerate(case["chunks"]):
            cache = O.TemporalCache()
            pos, parts = 0, []
            for n in chunks:
                parts.append(O.forward(w, cfg, px[:, pos:pos + n], cache=cache)["last_hidden_state"])
                pos += n
            got = sub(np.concatenate(parts, axis=1))
            _check(f"twin stream {chunks}", got, z[f"stream_{ci}_last_hidden_state_sub"])
            # streaming == one-shot
            _check(f"stream==full {chunks}", got, sub(one["last_hidden_state"]))


@pytest.mark.parametrize("name", FAST)
def test_oracle_matches_reference_golden(name):
    _run(name)


@pytest.mark.parametrize("name", SLOW)
def test_oracle_matches_reference_golden_full_depth(name):
    _run(name)


def test_oracle_causality_exact():
    """Perturbing the last frame must not change earlier frames at all."""
    cfg = O.OracleConfig(num_hidden_layers=1)
    w = O.make_weights(cfg, seed=3, style="stress")
    px = O.make_pixels(1, 4, cfg, seed=3)
    a = O.forward(w, cfg, px)["last_hidden_state"]
    px2 = px.copy()
    px2[:, -1] += 1.0
    b = O.forward(w, cfg, px2)["last_hidden_state"]
    assert np.array_equal(a[:, :-1], b[:, :-1])
    assert not np.array_equal(a[:, -1], b[:, -1])


def test_flops_accounting_matches_baseline_md():
    cfg = O.OracleConfig()
    assert abs(O.flops_per_clip(cfg, 16) / 16 / 1e9 - 49.40) < 0.02
    assert abs(O.flops_per_clip(cfg, 128) / 128 / 1e9 - 50.21) < 0.02
