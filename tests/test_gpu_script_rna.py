"""Whole-script parity for the RNA model: the call sequence of the unmodified reference script
2_GeneExpression/1_GeneExpress_train.py (tests/_rna_script.py - proved identical to the script itself on CPU by
tools/make_golden.py rna_script) runs on the GPU through the drop-in `models` module, resolved by bare name from dropin/
exactly like the script resolves it; what it hands to the concordance index at each of its 7 evaluate calls is compared
with the recording of the unmodified script on CPU (tests/golden/rna_script_reference.npz).

Tolerances: case order, survival months, vital status: bit-exact (grouping).  Scores: 1e-2 relative L2 (bf16 GEMMs,
fp32 accumulation - north_star's feature tolerance) in the deterministic variant (dropout_p = 0); with the script's
nn.Dropout() two random streams cannot agree better than ~10 % (tests/_rna_script.py), so that run is bounded at 0.3.
"""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dropin_models():
    """`from models import cox_loss, RNAOnlyModel` with dropin/ first on sys.path."""
    d = os.path.join(ROOT, "dropin")
    sys.modules.pop("models", None)
    sys.path.insert(0, d)
    try:
        return importlib.import_module("models")
    finally:
        sys.path.remove(d)


@pytest.mark.parametrize("fused_adam", [False, True])
def test_rna_train_script_sequence_matches_reference_recording(golden, fused_adam):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _rna_script as R
    from multimodalbrainsurvival_b200 import _lib, optim
    g = golden("rna_script_reference.npz")
    models = _dropin_models()
    hook = optim.accelerate_optimizer if fused_adam else None
    # ---- deterministic variant (dropout_p = 0): scores at 1e-2, TRAIN loss at 1e-3, head weights after two Adam steps
    l0 = _lib.launch_count()
    record, train_losses, last_state = R.run_like_script(models, torch.device("cuda:0"), optimizer_hook=hook, dropout_p=0.0)
    assert _lib.launch_count() > l0, "the drop-in did not launch a single libmmbs kernel"
    assert len(record) == int(g["n_calls"])
    for i, (months, neg_score, vital) in enumerate(record):
        assert np.array_equal(months, g[f"call{i}/months"]), f"call {i}: survival months / case order"
        assert np.array_equal(vital, g[f"call{i}/vital"]), f"call {i}: vital status / case order"
        ref = g[f"nodrop/call{i}/neg_score"].astype(np.float64)
        rel = np.linalg.norm(neg_score.astype(np.float64) - ref) / np.linalg.norm(ref)
        assert rel <= 1e-2, f"call {i}: scores differ from the reference's by {rel:.3g} (relative L2)"
    ref_losses = g["nodrop/train_losses"]
    assert np.all(np.abs(np.array(train_losses) - ref_losses) <= 1e-3 * np.abs(ref_losses)), (train_losses, ref_losses)
    w = last_state["final_mlp.0.weight"].numpy()
    assert np.abs(w - g["nodrop/final_head_weight"]).max() <= 2.5e-5, "head weights after two Adam steps (lr 1e-5)"
    # ---- the script's own configuration (nn.Dropout()): grouping bit-exact; scores only as close as two different
    # dropout streams allow (see tests/_rna_script.py: ~10 % after Adam's sign-like first steps)
    record, train_losses, _ = R.run_like_script(models, torch.device("cuda:0"), optimizer_hook=hook)
    for i, (months, neg_score, vital) in enumerate(record):
        assert np.array_equal(months, g[f"call{i}/months"]) and np.array_equal(vital, g[f"call{i}/vital"])
        ref = g[f"call{i}/neg_score"].astype(np.float64)
        rel = np.linalg.norm(neg_score.astype(np.float64) - ref) / np.linalg.norm(ref)
        assert np.isfinite(neg_score).all() and rel <= 0.3, f"call {i}: {rel:.3g}"
    ref_losses = g["train_losses"]
    assert np.all(np.abs(np.array(train_losses) - ref_losses) <= 0.15 * np.abs(ref_losses)), (train_losses, ref_losses)
