"""Shared helpers for the tests: load golden fixtures and rebuild their seeded inputs with the oracle."""
import json
import os

import numpy as np

from oracle import streamformer_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUB_TOK, SUB_DIM = 14, 8


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = json.loads(str(z["case"]))
    return case, z


def case_config(case) -> O.OracleConfig:
    return O.OracleConfig(num_hidden_layers=case["layers"], enable_causal_temporal=case["causal"],
                          add_lora_spatial=case["lora"], num_frames=case.get("num_frames", 16))


def case_inputs(case):
    cfg = case_config(case)
    w = O.make_weights(cfg, seed=case["seed"], style=case["style"])
    px = O.make_pixels(case["B"], case["T"], cfg, seed=case["seed"], H=case.get("H"), W=case.get("W"))
    return cfg, w, px


def sub(x):
    return np.ascontiguousarray(np.asarray(x)[..., ::SUB_TOK, ::SUB_DIM])


def rel_rms(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / max(np.sqrt((b ** 2).mean()), 1e-30))


def cosine(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-30))
