"""GPU parity of the fused clip + SGD + EMA step (csrc/optim.cu) against the golden trajectory produced by the
REFERENCE's ModelEMA + scheduler with torch.optim.SGD and clip_grad_norm_ (tests/golden/step_sgd_ema.npz).

Tolerance (floating point, stated): 2e-6 relative per step on parameters and EMA (the forward/backward that feeds
the gradients runs in torch on the GPU with a different summation order than the CPU golden)."""
import copy
import importlib.util
import os

import numpy as np
import pytest
import torch

from visiondk_b200.optim import FusedSGDClipEMA

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def flat_state(m):
    return np.concatenate([p.detach().cpu().numpy().ravel().astype(np.float32) for p in m.state_dict().values()])


def test_fused_step_follows_reference_trajectory(lib):
    z = np.load(os.path.join(GOLD, "step_sgd_ema.npz"))
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3))
    assert np.array_equal(flat_state(model), z["init"])  # same seeded init as the golden run
    model = model.cuda()
    ema_model = copy.deepcopy(model).eval()
    for p in ema_model.parameters():
        p.requires_grad_(False)
    groups = [{"params": model[0].parameters(), "lr": 0.01},
              {"params": list(model[1].parameters()) + list(model[2].parameters()), "lr": 0.1}]
    opt = FusedSGDClipEMA(groups, lr=0.01, momentum=0.8, weight_decay=5e-4, max_norm=10.0, model=model, ema_model=ema_model)
    for step in range(z["x"].shape[0]):
        for pg, lr in zip(opt.param_groups, z["lrs"][step]):
            pg["lr"] = float(lr)  # the golden lrs are those in effect during this step
        x = torch.from_numpy(z["x"][step]).cuda()
        y = torch.from_numpy(z["y"][step]).cuda()
        loss = torch.nn.functional.cross_entropy(model(x) * 30, y)
        loss.backward()
        opt.step()
        assert abs(opt.grad_norm() - z["gnorms"][step]) <= 1e-4 * z["gnorms"][step]
        np.testing.assert_allclose(flat_state(model), z["params"][step], rtol=3e-5, atol=3e-6, err_msg=f"params step {step}")
        np.testing.assert_allclose(flat_state(ema_model), z["emas"][step], rtol=3e-5, atol=3e-6, err_msg=f"ema step {step}")
        assert all((p.grad == 0).all() for p in model.parameters())


def test_fused_step_matches_torch_sgd_on_large_flat_buffers(lib):
    torch.manual_seed(0)
    n = 3_000_001
    p0 = torch.randn(n, device="cuda")
    g0 = torch.randn(n, device="cuda") * 0.01
    ref_p = torch.nn.Parameter(p0.clone())
    ref = torch.optim.SGD([ref_p], lr=0.05, momentum=0.9, weight_decay=1e-3)
    ours_p = torch.nn.Parameter(p0.clone())
    opt = FusedSGDClipEMA([ours_p], lr=0.05, momentum=0.9, weight_decay=1e-3, max_norm=10.0)
    for it in range(3):
        ref_p.grad = g0.clone() * (it + 1)
        ours_p.grad.copy_(g0 * (it + 1))
        torch.nn.utils.clip_grad_norm_([ref_p], 10.0)
        ref.step()
        opt.step()
        torch.testing.assert_close(ours_p.data, ref_p.data, rtol=2e-6, atol=2e-7)
