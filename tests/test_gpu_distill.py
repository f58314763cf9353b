"""GPU parity of the distillation objectives against fixtures of the reference itself
(tests/golden/kd_*.npz, oracle/make_golden_r2.py): the fused path (`FusedTrainStep(distillation=...)`:
ncf_loss_grad_kd, ncf_backward, ncf_feature_kd) and the autograd path (the reference's own loop
`loss = distillation(u, i, y); loss.backward(); optimizer.step()`, scripts/train_student.py:148-156)."""
import numpy as np
import pytest
import torch

from tests.util import assert_close, assert_close_adam, group, load_golden

pytestmark = pytest.mark.gpu
CASES = ["kd_response_cfg3", "kd_soft_target", "kd_feature", "kd_feature_same", "kd_attention"]


def dev():
    return torch.device("cuda:0")


def build(name):
    from ncf_b200 import distillation as D
    from ncf_b200.models import NCF
    z, meta = load_golden(name)

    def model(prefix, cfg):
        m = NCF(meta["U"], meta["I"], cfg["f"], cfg["L"], 0.0, "NeuMF-end")
        m.load_state_dict({k: torch.from_numpy(v) for k, v in group(z, prefix).items()})
        return m.to(dev())
    teacher, student = model("teacher", meta["teacher"]), model("init", meta["student"])
    kind, a, T = meta["kind"], meta["alpha"], meta["temperature"]
    if kind == "response":
        dist = D.ResponseDistillation(teacher, student, temperature=T, alpha=a)
    elif kind == "soft":
        dist = D.SoftTargetDistillation(teacher, student, temperature=T, alpha=a)
    elif kind == "feature":
        dist = D.FeatureDistillation(teacher, student, temperature=T, alpha=a, beta=meta["beta"])
        with torch.no_grad():   # the reference's adapters are random: take the very ones the fixture ran with
            for key, lin in dist.adaptation_layers.items():
                lin.weight.copy_(torch.from_numpy(z[f"adapter/{key}.weight"]))
                lin.bias.copy_(torch.from_numpy(z[f"adapter/{key}.bias"]))
        assert set(dist.adaptation_layers.keys()) == {k.split("/")[1].split(".")[0] for k in z.files if k.startswith("adapter/")}
    else:
        dist = D.AttentionDistillation(teacher, student, temperature=T, alpha=a, gamma=meta["gamma"])
    return z, meta, dist.to(dev()), student


def batch(z, t):
    return tuple(torch.from_numpy(z[k][t]).to(dev()) for k in ("user", "item", "label"))


@pytest.mark.parametrize("name", CASES)
def test_autograd_path_matches_reference(name):
    """The reference loop, unchanged, on the drop-in classes."""
    z, meta, dist, student = build(name)
    opt = torch.optim.Adam(student.parameters(), lr=meta["lr"])
    for t in range(meta["T"]):
        dist.train()
        u, i, y = batch(z, t)
        opt.zero_grad()
        loss = dist(u, i, y)
        loss.backward()
        assert abs(loss.item() - z["loss"][t]) <= 5e-6 * abs(z["loss"][t]), (t, loss.item(), z["loss"][t])
        if t == 0:
            for k, want in group(z, "grad0").items():
                got = dict(student.named_parameters())[k].grad.cpu().numpy()
                assert_close(got, want, f"grad {k}")
        opt.step()
    for k, want in group(z, "final").items():
        assert_close_adam(student.state_dict()[k].cpu().numpy(), want, f"final {k}")


@pytest.mark.parametrize("name", [c for c in CASES if c != "kd_feature_same"])
@pytest.mark.parametrize("dense", ["0", "1"])
def test_fused_path_matches_reference(monkeypatch, name, dense):
    """FusedTrainStep(distillation=...): no autograd, native loss / feature-matching kernels, lazy or
    all-rows Adam."""
    from ncf_b200.trainer import FusedTrainStep
    monkeypatch.setenv("NCF_ADAM_DENSE", dense)
    z, meta, dist, student = build(name)
    ts = FusedTrainStep(student, "adam", meta["lr"], max_batch=meta["B"], distillation=dist)
    for t in range(meta["T"]):
        ts.step(*batch(z, t))
        loss = ts.pop_loss()
        want = z["loss"][t]
        # the attention-transfer term (~1e-9 * gamma) is left out of the fused objective
        assert abs(loss - want) <= 5e-6 * abs(want), (t, loss, want)
    ts.flush()
    for k, want in group(z, "final").items():
        assert_close_adam(student.state_dict()[k].cpu().numpy(), want, f"final {k}")


def test_fused_path_refuses_tower_level_features():
    """Equal architectures: mlp_linear_k / mlp_relu_k match too (feature.py:71-79); the fused path says so
    instead of silently dropping them."""
    from ncf_b200.trainer import FusedTrainStep
    z, meta, dist, student = build("kd_feature_same")
    with pytest.raises(NotImplementedError):
        FusedTrainStep(student, "adam", 1e-3, max_batch=meta["B"], distillation=dist)


def test_fused_gradients_of_the_feature_objective_match_oracle():
    """ncf_loss_grad_kd + ncf_backward + ncf_feature_kd gradient buffers against the oracle's
    feature_matching / backward (first batch of the fixture)."""
    from ncf_b200 import ops
    from ncf_b200.trainer import FusedTrainStep
    from oracle import ncf_numpy as onp
    z, meta, dist, student = build("kd_feature")
    ts = FusedTrainStep(student, "adam", meta["lr"], max_batch=meta["B"], distillation=dist)
    u, i, y = batch(z, 0)
    ops.mark_rows(ts._m, ts._g, u, i)
    t_logits = ts.teacher_logits[:u.numel()]
    ops.forward(ts._tm, u, i, out=t_logits, workspace=ts.teacher_workspace)
    ts._kd_grads(u, i, y, t_logits)
    torch.cuda.synchronize()
    want = group(z, "grad0")
    g = ts.grads
    for k, buf in (("embed_user_GMF.weight", g.g_user_gmf), ("embed_item_GMF.weight", g.g_item_gmf),
                   ("embed_user_MLP.weight", g.g_user_mlp), ("embed_item_MLP.weight", g.g_item_mlp)):
        assert_close(buf.cpu().numpy(), want[k], f"grad {k}")
    # and the oracle agrees with what was injected
    student_np, teacher_np = group(z, "init"), group(z, "teacher")
    adapters = {key: (z[f"adapter/{key}.weight"], z[f"adapter/{key}.bias"]) for key in ("gmf_features", "mlp_input")}
    _, dfeat = onp.feature_matching(student_np, teacher_np, z["user"][0], z["item"][0], adapters, meta["beta"])
    assert set(dfeat) == {"gmf_features", "mlp_input"}


def test_soft_target_helper_matches_oracle():
    """BaseDistillation.knowledge_distillation_loss / task_loss as standalone calls (base.py:26-38)."""
    from ncf_b200 import distillation as D
    from oracle import ncf_numpy as onp
    z, meta, dist, student = build("kd_soft_target")
    rng = np.random.default_rng(0)
    x, t = rng.standard_normal(257).astype(np.float32) * 3, rng.standard_normal(257).astype(np.float32) * 3
    y = (rng.random(257) < 0.4).astype(np.float32)
    xd, td, yd = (torch.from_numpy(a).to(dev()) for a in (x, t, y))
    xd.requires_grad_(True)
    kd = D.BaseDistillation.knowledge_distillation_loss(dist, td, xd)
    kd.backward()
    want, dl = onp.kd_loss_and_dlogit(x, y, t, 0.0, 1.0, dist.temperature, 1)
    assert abs(kd.item() - float(want)) <= 5e-6 * abs(float(want))
    assert_close(xd.grad.cpu().numpy(), dl, "dKD/dx")
    task = dist.task_loss(xd.detach(), yd)
    assert abs(task.item() - float(onp.bce_with_logits(x, y).mean())) <= 5e-6
