"""Config C1 (single-qubit phase): the native notebook MLPs / BitstringDDM against fixtures produced by the notebook's
own classes (tests/golden/nb_mlp.npz, generated from NB c6 / c12) and against oracle autograd."""
import numpy as np
import pytest
import torch

from conftest import golden_state_dict, load_golden
from oracle import ddqst_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cls", ["SimpleMLP", "UpgradedMLP"])
def test_notebook_mlp_forward_sample_and_gradients(cls):
    import ddqst_b200 as dq
    z = load_golden("nb_mlp.npz")
    T = int(z["T"][0])
    shots, basis, seed, off = (int(v) for v in z["sample_args"])
    sd = golden_state_dict(z, f"{cls}.sd.")
    model = getattr(dq, cls)(T, 3)
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd)
    ddm = dq.BitstringDDM(model, T, "cuda", seed=seed)
    x, t, b = (torch.from_numpy(z[f"{cls}.{k}"]).cuda() for k in ("x", "t", "basis"))
    with torch.no_grad():
        logits = model(x, t, b)
    assert np.abs(logits.cpu().numpy() - z[f"{cls}.logits"]).max() < 1e-4                       # vs the notebook class itself
    # sampler: fp32 logits differ from the CPU run by ~1e-6, a draw flips only inside that margin
    got = ddm.sample(shots, basis, shot_offset=off)
    assert got.shape == (shots,) and (got == z[f"{cls}.samples"]).mean() >= 0.99
    # forward_diffusion is bit exact under the injected stream
    _, Q = orc.notebook_schedule(T)
    x0 = torch.randint(0, 2, (5000,))
    tt = torch.randint(1, T + 1, (5000,))
    want = orc.q_sample_marginal(Q, x0.view(-1, 1), tt, seed, 9, row_offset=3)[:, 0]
    assert torch.equal(ddm.forward_diffusion(x0, tt, stream_id=9, row_offset=3).cpu(), want)
    # the notebook's training loop body: loss = ddm.train_step(...); loss.backward(); optimizer.step()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    g = torch.Generator().manual_seed(1)
    xt = torch.randint(0, 2, (200,), generator=g)
    x0 = torch.randint(0, 2, (200,), generator=g)
    t2 = torch.randint(1, T + 1, (200,), generator=g)
    b2 = torch.randint(0, 3, (200,), generator=g)
    want_loss = torch.nn.functional.cross_entropy(orc.notebook_mlp_forward(params, xt, t2, b2), x0)
    want_loss.backward()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss = torch.nn.functional.cross_entropy(model(xt.cuda(), t2.cuda(), b2.cuda()), x0.cuda())
    opt.zero_grad()
    loss.backward()
    assert abs(loss.item() - want_loss.item()) < 1e-5
    for name, p in model.named_parameters():
        assert torch.allclose(p.grad.cpu(), params[name].grad, atol=1e-6), name
    opt.step()
    l2 = ddm.train_step(x0, b2)
    assert torch.isfinite(l2) and l2.requires_grad
