"""GPU parity of csrc/nll.cu (NLLSurvLoss / nll_loss, the `survival_bin` task) against the
reference's own nll_loss + autograd (tests/golden/nll_reference.npz, tools/make_golden.py nll).  fp32, tolerance 1e-5
relative on the loss and on the gradient (relative to its largest entry)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_nll_surv_loss_matches_reference(golden):
    from multimodalbrainsurvival_b200 import _lib, models
    g = golden("nll_reference.npz")
    for ci in range(4):
        h = torch.tensor(g[f"case{ci}/h"], device="cuda:0", requires_grad=True)
        y = torch.tensor(g[f"case{ci}/y"], device="cuda:0")
        c = torch.tensor(g[f"case{ci}/c"], device="cuda:0")
        crit = models.NLLSurvLoss(alpha=float(g[f"case{ci}/alpha"]), reduction="mean" if int(g[f"case{ci}/mean"]) else "sum")
        l0 = _lib.launch_count()
        loss = crit(h, y, c)
        (loss * 1.5).backward()
        assert _lib.launch_count() >= l0 + 3, "the loss did not run on the libmmbs kernels"
        ref_loss, ref_grad = float(g[f"case{ci}/loss"]), g[f"case{ci}/grad"] * 1.5
        assert abs(float(loss) - ref_loss) <= 1e-5 * abs(ref_loss) + 1e-7, (ci, float(loss), ref_loss)
        err = np.abs(h.grad.cpu().numpy() - ref_grad).max()
        assert err <= 1e-5 * np.abs(ref_grad).max() + 1e-9, (ci, err)


def test_nll_out_of_range_bin_is_flagged_not_silently_wrong():
    from multimodalbrainsurvival_b200 import models
    h = torch.zeros(3, 4, device="cuda:0")
    y = torch.tensor([0, 7, 1], device="cuda:0")
    c = torch.zeros(3, device="cuda:0")
    assert torch.isnan(models.nll_loss(h, y.view(3, 1), c.view(3, 1)))
