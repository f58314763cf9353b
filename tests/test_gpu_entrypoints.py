"""End-to-end runs of the four entry points on a tiny synthetic shape (GPU): CLI contract, stdout
RESULTS block, checkpoint names/layout, learning signal (HR well above chance), pretrain ->
NeuMF-pre and teacher -> student chains."""
import re
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "scripts"))


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    from ncf_b200.config import config
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(config, "output_dir", tmp_path / "results")
    monkeypatch.setattr(config, "model_dir", tmp_path / "results" / "models")
    monkeypatch.setattr(config, "log_dir", tmp_path / "results" / "logs")
    monkeypatch.setattr(config, "figure_dir", tmp_path / "results" / "figures")
    return tmp_path


def parse_results(out):
    block = out.split("--- RESULTS ---")[1].split("--- END RESULTS ---")[0]
    return {k.strip(): v.strip() for k, v in (l.split(":", 1) for l in block.strip().splitlines())}


def test_pretrain_then_neumf_pre_then_distil(workdir, capsys):
    import pretrain
    import train_neumf
    import train_student
    import train_teacher
    common = ["--synthetic", "tiny", "--factor_num", "8", "--batch_size", "256", "--lr", "0.005"]
    pretrain.main(["--model", "GMF", "--epochs", "6", "--num_layers", "3", *common])
    out = capsys.readouterr().out
    r = parse_results(out)
    assert set(r) == {"HR@10", "NDCG@10", "Parameters"}
    assert re.search(r"Epoch 001: Loss=\d\.\d{4}, HR=\d\.\d{3}, NDCG=\d\.\d{3}, Time=", out)
    assert float(r["HR@10"]) > 0.2  # chance is 0.10 with 99 negatives
    pretrain.main(["--model", "MLP", "--epochs", "6", "--num_layers", "3", *common])
    capsys.readouterr()
    models = workdir / "results" / "models"
    assert (models / "GMF_8f_best.pth").exists() and (models / "MLP_3l_8f_best.pth").exists()

    res = train_neumf.main(["--model", "NeuMF-pre", "--pretraining", "--epochs", "3", "--num_layers", "3",
                            "--synthetic", "tiny", "--factor_num", "8", "--batch_size", "256", "--lr", "0.001"])
    out = capsys.readouterr().out
    r = parse_results(out)
    assert r["Model"] == "NeuMF-pre" and r["Pretraining"] == "True" and r["Layers"] == "3"
    assert int(r["Parameters"]) == res["parameters"]
    ckpt = torch.load(models / "NeuMF_pre_3l_8f_best.pth")
    assert list(ckpt)[0] == "embed_user_GMF.weight" and list(ckpt)[-1] == "predict_layer.bias"
    assert res["best_hr"] > 0.2

    # teacher (f=16, L=3) -> student (f=8, L=2): the reference's (2f, L+1) rule
    train_teacher.main(["--model", "NeuMF-end", "--epochs", "5", "--num_layers", "3", "--synthetic", "tiny",
                        "--factor_num", "16", "--batch_size", "256", "--lr", "0.005"])
    capsys.readouterr()
    assert (models / "teacher_NeuMF-end_best.pth").exists()
    st = train_student.main(["--teacher_model", "NeuMF-end", "--epochs", "5", "--num_layers", "2",
                             "--factor_num", "8", "--synthetic", "tiny", "--batch_size", "256", "--lr", "0.005"])
    out = capsys.readouterr().out
    assert re.search(r"000 - Loss: \d+\.\d{6}, HR: \d\.\d{3}, NDCG: \d\.\d{3}, Time: \d\d:\d\d:\d\d", out)
    assert (models / "student_NeuMF-end_best.pth").exists()
    assert st.best_hr > 0.15
    # the other strategies of reference scripts/train_student.py:96-127 run on the fused path as well
    for strategy in ("feature", "attention", "unified"):
        st = train_student.main(["--teacher_model", "NeuMF-end", "--epochs", "3", "--num_layers", "2", "--factor_num", "8",
                                 "--synthetic", "tiny", "--batch_size", "256", "--lr", "0.005", "--distillation", strategy])
        capsys.readouterr()
        assert st.best_hr > 0.15, strategy


def test_tf32_tower_mode_within_stated_tolerance():
    """tower_math='tf32' (single-pass TF32): logits within 2e-3 of the fp32-parity mode relative to
    the logit scale — the stated tolerance for this opt-in mode."""
    from ncf_b200.models import NCF
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = NCF(500, 400, 32, 3, 0.0, "NeuMF-end").to(dev).eval()
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k.startswith("embed_"):
                p.mul_(20.0)
    u = torch.randint(0, 500, (4096,), device=dev)
    i = torch.randint(0, 400, (4096,), device=dev)
    with torch.no_grad():
        ref = model(u, i)
        model.tower_math = "tf32"
        fast = model(u, i)
    err = (fast - ref).abs().max().item() / ref.abs().max().item()
    assert 0 < err < 2e-3, err


@pytest.mark.parametrize("depth", [0, 2, 3])
def test_host_fed_trainer_equals_eager_steps(depth):
    """HostFedTrainer (graph-captured steps fed from pinned host batches through a copy stream)
    follows the same trajectory as eager FusedTrainStep.step on device batches — also with two steps in
    flight (the loss of step k read while step k+1 runs)."""
    import copy
    from ncf_b200.models import NCF
    from ncf_b200.trainer import FusedTrainStep, HostFedTrainer
    from tests.util import assert_close_adam
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    U, I, B, T = 300, 200, 512, 5
    a = NCF(U, I, 8, 3, 0.0, "NeuMF-end").to(dev)
    b = copy.deepcopy(a)
    g = torch.Generator().manual_seed(4)
    batches = [(torch.randint(0, U, (B,), generator=g).pin_memory(), torch.randint(0, I, (B,), generator=g).pin_memory(),
                (torch.rand(B, generator=g) < 0.3).float().pin_memory()) for _ in range(T)]
    ta = FusedTrainStep(a, "adam", 1e-3, max_batch=B)
    tb = FusedTrainStep(b, "adam", 1e-3, max_batch=B)
    eager = []
    for u, i, y in batches:
        ta.step(u.to(dev), i.to(dev), y.to(dev))
        eager.append(ta.pop_loss())
    pipelined = depth > 0
    hf = HostFedTrainer(tb, B, depth=max(depth, 2))
    hf.prefetch(*batches[0])
    fed = []
    for k in range(T):
        hf.launch()
        if k + 1 < T:
            hf.prefetch(*batches[k + 1])
        if not pipelined or hf.in_flight == hf.depth:
            fed.append(hf.wait())
    while hf.in_flight:
        fed.append(hf.wait())
    with pytest.raises(Exception):
        hf.wait()                      # nothing in flight any more
    assert len(fed) == T
    for x, y in zip(eager, fed):
        assert abs(x - y) <= 1e-6 * abs(x)
    ta.flush(); tb.flush()
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        assert_close_adam(sb[k].cpu().numpy(), sa[k].cpu().numpy(), k)
