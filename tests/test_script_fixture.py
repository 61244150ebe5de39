"""CPU check of the whole-script fixture (tests/golden/rna_script_reference.npz, written by tools/make_golden.py
rna_script from the UNMODIFIED reference script): its evaluate calls cover the synthetic splits of tests/_rna_script.py
in the script's order (train, val per epoch; val, val, test at the end) and carry one score per case."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _rna_script as R  # noqa: E402


def test_recording_matches_the_synthetic_splits(golden):
    g = golden("rna_script_reference.npz")
    order = ["train", "val"] * R.CONFIG["num_epochs"] + ["val", "val", "test"]
    assert int(g["n_calls"]) == len(order)
    for i, split in enumerate(order):
        cases, months, vital, _ = R.synthetic_split(split)
        idx = np.argsort(cases, kind="stable")          # get_survival_CI: sorted(set(ids))
        assert np.array_equal(g[f"call{i}/months"], months[idx]), (i, split)
        assert np.array_equal(g[f"call{i}/vital"], vital[idx]), (i, split)
        assert g[f"call{i}/neg_score"].shape == (len(cases),) and np.isfinite(g[f"call{i}/neg_score"]).all()
    assert np.isfinite(g["train_losses"]).all() and len(g["train_losses"]) == R.CONFIG["num_epochs"]
