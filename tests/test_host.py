"""Host-side logic that runs without a GPU: the drop-in module surface (constructor, attribute
names, state_dict layout, initialisation, NeuMF-pre loading), data file parsing, config defaults,
and the multi-GPU plan (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.nn as nn

from tests.util import group, load_golden


def test_state_dict_layout_matches_reference():
    """Keys, order, shapes and dtype of the checkpoint."""
    from ncf_b200.models import NCF
    z, meta = load_golden("train_neumf_f8_l3")
    ref = group(z, "init")
    ref_keys = [k[len("init/"):] for k in z.files if k.startswith("init/")]
    m = NCF(meta["U"], meta["I"], meta["f"], meta["L"], 0.0, "NeuMF-end")
    sd = m.state_dict()
    assert list(sd.keys()) == ref_keys
    for k in ref_keys:
        assert tuple(sd[k].shape) == ref[k].shape and sd[k].dtype == torch.float32
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ref.items()})  # strict load of a reference ckpt
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == sum(v.size for v in ref.values())


def test_same_seed_gives_reference_initial_weights():
    """torch.manual_seed(s); NCF(...) draws the same numbers in the same order as the reference
    (src/ncf/models.py:11-46)."""
    from ncf_b200.models import NCF
    z, meta = load_golden("init_seed2025")
    torch.manual_seed(meta["seed"])
    m = NCF(meta["U"], meta["I"], meta["f"], meta["L"], 0.0, meta["model_type"])
    for k, ref in group(z, "init").items():
        assert np.array_equal(m.state_dict()[k].numpy(), ref), k


@pytest.mark.parametrize("mt,ps", [("GMF", 8), ("MLP", 8), ("NeuMF-end", 16), ("NeuMF-pre", 16)])
def test_module_surface(mt, ps):
    from ncf_b200.models import NCF
    m = NCF(30, 20, 8, 3, 0.0, mt)
    assert m.model_type == mt and m.dropout == 0.0
    for name, rows, dim in (("embed_user_GMF", 30, 8), ("embed_item_GMF", 20, 8),
                            ("embed_user_MLP", 30, 32), ("embed_item_MLP", 20, 32)):
        emb = getattr(m, name)
        assert isinstance(emb, nn.Embedding) and emb.embedding_dim == dim and emb.weight.shape == (rows, dim)
    kinds = [type(x) for x in m.MLP_layers]
    assert kinds == [nn.Dropout, nn.Linear, nn.ReLU] * 3
    assert [(l.in_features, l.out_features) for l in m.linears()] == [(64, 32), (32, 16), (16, 8)]
    assert m.predict_layer.in_features == ps and m.predict_layer.out_features == 1
    with pytest.raises(ValueError):
        NCF(30, 20, 8, 3, 0.0, "bogus")


def test_load_pretrain_weights_semantics():
    from ncf_b200.models import NCF
    z, meta = load_golden("neumf_pre_sgd")
    gmf = {k: torch.from_numpy(v) for k, v in group(z, "gmf").items()}
    mlp = {k: torch.from_numpy(v) for k, v in group(z, "mlp").items()}
    m = NCF(meta["U"], meta["I"], meta["f"], meta["L"], 0.0, "NeuMF-pre")
    torch.manual_seed(meta["reseed"])
    m.load_pretrain_weights(gmf, mlp)
    for k, ref in group(z, "init").items():
        assert np.array_equal(m.state_dict()[k].numpy(), ref), k
    assert float(m.predict_layer.bias.abs().sum()) == 0.0
    # a NeuMF-end model ignores the call (reference models.py:50-51)
    e = NCF(meta["U"], meta["I"], meta["f"], meta["L"], 0.0, "NeuMF-end")
    before = {k: v.clone() for k, v in e.state_dict().items()}
    e.load_pretrain_weights(gmf, mlp)
    assert all(torch.equal(before[k], v) for k, v in e.state_dict().items())
    # a broken checkpoint raises RuntimeError like the reference (models.py:91-95)
    with pytest.raises(RuntimeError):
        m.load_pretrain_weights({}, mlp)


def test_dropout_in_training_mode_is_refused():
    from ncf_b200.models import NCF
    m = NCF(5, 5, 8, 2, 0.5, "NeuMF-end").train()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, dtype=torch.int64), torch.zeros(1, dtype=torch.int64))


def test_data_files_round_trip(tmp_path, monkeypatch):
    """u.train.rating / u.test.negative in the reference's format (preprocessing.py:137-154) parse
    into the structures load_all() returns (datasets.py:9-36)."""
    from ncf_b200 import datasets
    from ncf_b200.synth import make_interactions
    inter = make_interactions((50, 160, 1500), device="cpu", n_test_neg=99, seed=3)
    tr, te = tmp_path / "u.train.rating", tmp_path / "u.test.negative"
    datasets.write_reference_files(inter, tr, te)
    assert not te.read_text().endswith("\n")
    monkeypatch.setattr(datasets.config, "train_rating", tr)
    monkeypatch.setattr(datasets.config, "test_negative", te)
    train_data, test_data, user_num, item_num, train_mat = datasets.load_all(host=True)
    assert user_num == int(inter.pos_user.max()) + 1 and item_num == int(inter.pos_item.max()) + 1
    assert np.array_equal(train_data, torch.stack([inter.pos_user, inter.pos_item], 1).numpy())
    assert test_data.shape == (50 * 100, 2)
    assert np.array_equal(test_data[:, 1].reshape(50, 100), inter.test_cands.numpy())
    assert np.array_equal(test_data[::100, 0], inter.test_users.numpy())
    u0, i0 = int(train_data[0, 0]), int(train_data[0, 1])
    assert (u0, i0) in train_mat and (u0, int(inter.test_cands[u0, 1])) not in train_mat
    assert len(train_mat.nonzero()[0]) == len(train_data)
    ds = datasets.NCFData(test_data, item_num, train_mat, 0, False)
    assert len(ds) == 5000 and ds[100] == (int(test_data[100, 0]), int(test_data[100, 1]), 0)
    with pytest.raises(AssertionError):
        ds.ng_sample()  # "no need to sampling when testing"


def test_synthetic_data_is_leak_free_and_sorted():
    from ncf_b200.synth import make_interactions
    d = make_interactions((80, 300, 4000), device="cpu", seed=1)
    seen = set(zip(d.pos_user.tolist(), d.pos_item.tolist()))
    assert len(seen) == d.pos_user.numel()
    for u, row in zip(d.test_users.tolist(), d.test_cands.tolist()):
        assert all((u, c) not in seen for c in row)       # held-out item and negatives unseen in train
        assert row[1:] == sorted(row[1:]) and len(set(row)) == 100
    again = make_interactions((80, 300, 4000), device="cpu", seed=1)
    assert torch.equal(d.pos_item, again.pos_item)            # seeded


def test_config_defaults_are_the_reference_yaml_values():
    from ncf_b200.config import Config
    c = Config("/nonexistent.yaml")
    assert (c.factor_num, c.num_layers, c.dropout, c.batch_size, c.epochs, c.lr, c.num_ng,
            c.test_num_ng, c.top_k, c.temperature, c.alpha, c.user_num, c.item_num) == \
           (32, 2, 0.0, 256, 20, 0.001, 4, 99, 10, 2.0, 0.5, 944, 1683)
    assert str(c.model_dir) == "results/models"


def test_torch_port_agrees_with_numpy_oracle():
    from oracle import ncf_numpy as onp
    from oracle import torch_port as tp
    P = tp.init_params(40, 30, 8, 3, "NeuMF-end", seed=1)
    params = {k: v.detach().numpy().copy() for k, v in P.items()}
    rng = np.random.default_rng(0)
    u, i = rng.integers(0, 40, 64), rng.integers(0, 30, 64)
    y = (rng.random(64) < 0.3).astype(np.float32)
    tr = tp.CpuTrainer(P, "NeuMF-end")
    opt = onp.DenseAdam()
    for _ in range(3):
        logits = onp.forward(params, u, i, "NeuMF-end")
        loss, dl = onp.loss_and_dlogit(logits, y)
        opt.step(params, onp.backward(params, u, i, "NeuMF-end", dl))
        got = tr.step(torch.from_numpy(u), torch.from_numpy(i), torch.from_numpy(y))
        assert abs(got - float(loss)) < 1e-5
    for k in params:
        err = np.abs(P[k].detach().numpy() - params[k]).max() / max(np.abs(params[k]).max(), 1e-30)
        assert err < 1e-4, k


# ---- multi-GPU plan on CPU (gloo, world_size 2) ----------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from ncf_b200 import dist as nd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        B = 8
        user = torch.randint(0, 50, (world * B,), generator=g)
        item = torch.randint(0, 40, (world * B,), generator=g)
        lo, hi = nd.partition(world * B, world, rank)
        gu, gi = nd.gather_indices(user[lo:hi].contiguous(), item[lo:hi].contiguous(), world)
        ok = torch.equal(gu, user) and torch.equal(gi, item)
        # averaging per-rank mean gradients == gradient of the global mean
        local = torch.full((5,), float(rank + 1))
        nd.average_(local, world)
        ok = ok and torch.allclose(local, torch.full((5,), (1 + world) / 2))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_replicated_dp_plan_gloo_world2():
    import torch.multiprocessing as mp
    from ncf_b200 import dist as nd
    assert [nd.partition(10, 3, r) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] and out[1]


def test_reference_import_paths_resolve_to_this_package():
    """`from src.ncf.models import NCF` etc. (the reference's import paths) work unchanged."""
    from src.data.datasets import NCFData, load_all  # noqa: F401
    from src.distillation import ResponseDistillation
    from src.ncf.models import NCF
    from src.training.metrics import metrics
    from src.utils.config import config
    assert NCF.__module__ == "ncf_b200.models" and metrics.__module__ == "ncf_b200.metrics"
    assert ResponseDistillation.__module__ == "ncf_b200.distillation" and config.batch_size == 256


def test_row_shard_layout_helpers():
    """Row r -> rank r % world at local index r // world; shards tile the table exactly."""
    from ncf_b200.dist import shard_rows, shard_state_dict
    assert [shard_rows(10, 4, r) for r in range(4)] == [3, 3, 2, 2]
    assert sum(shard_rows(301, 8, r) for r in range(8)) == 301
    full = {"embed_user_GMF.weight": torch.arange(14.0).reshape(7, 2), "predict_layer.bias": torch.ones(1)}
    parts = [shard_state_dict(full, 3, r) for r in range(3)]
    assert parts[1]["embed_user_GMF.weight"][:, 0].tolist() == [2.0, 8.0]      # rows 1 and 4
    assert all(torch.equal(p["predict_layer.bias"], full["predict_layer.bias"]) for p in parts)
    rebuilt = torch.empty(7, 2)
    for r in range(3):
        rebuilt[r::3] = parts[r]["embed_user_GMF.weight"]
    assert torch.equal(rebuilt, full["embed_user_GMF.weight"])


def test_optimiser_mode_heuristic():
    """FusedTrainStep.dense_adam: all-rows Adam only when a step is expected to touch a large share of
    the tables (the decision is host logic; it must not need a GPU)."""
    from ncf_b200.trainer import FusedTrainStep

    class _M:
        user_num, item_num = 138_493, 26_744

    ts = FusedTrainStep.__new__(FusedTrainStep)
    ts.model, ts.dense_share = _M(), 0.42
    assert ts.dense_adam(65_536)            # the bench batch: ~46 % of the rows distinct
    assert ts.dense_adam(8 * 65_536)        # an 8-GPU global batch touches nearly everything
    assert not ts.dense_adam(256)           # the reference batch size
    _M.user_num, _M.item_num = 10_000_000, 1_000_000
    assert not ts.dense_adam(65_536)        # 10M x 1M tables: < 1 % of the rows per step
    import os
    os.environ["NCF_ADAM_DENSE"] = "1"
    try:
        assert ts.dense_adam(1)
    finally:
        del os.environ["NCF_ADAM_DENSE"]


def test_distillation_fused_specs_describe_the_reference_objectives():
    """What FusedTrainStep is told to run for each strategy (no device work): weights of the task / KD terms as
    reference base.py:40-50, response.py:43-61, feature.py:138-146, attention.py:93-101 combine them, which
    embedding-level features are matched (adapters where widths differ) and what is refused."""
    from ncf_b200.distillation import (AttentionDistillation, FeatureDistillation, ResponseDistillation,
                                       SoftTargetDistillation, UnifiedDistillation)
    from ncf_b200.models import NCF
    torch.manual_seed(0)
    teacher = NCF(30, 20, 16, 4, 0.0, "NeuMF-end")     # the reference script's rule: (2f, L+1) of the student
    student = NCF(30, 20, 8, 3, 0.0, "NeuMF-end")
    r = ResponseDistillation(teacher, student, 2.0, 0.5).fused_spec()
    assert (r["w_task"], r["w_kd"], r["kd_mode"], r["features"]) == (0.5, 0.5, 0, [])
    s = SoftTargetDistillation(teacher, student).fused_spec()          # defaults T=4, alpha=0.7
    assert s["kd_mode"] == 1 and s["temperature"] == 4.0 and abs(s["w_task"] - 0.7) < 1e-12 and abs(s["w_kd"] - 0.3) < 1e-12
    fd = FeatureDistillation(teacher, student, 2.0, 0.5, 0.3)
    assert sorted(fd.matched_keys()) == ["gmf_features", "mlp_input"]  # tower widths differ: skipped, as in the reference
    f = fd.fused_spec()
    assert abs(f["w_kd"] - 0.2) < 1e-12 and [x["kind"] for x in f["features"]] == [0, 1]
    assert all(abs(x["weight"] - 0.15) < 1e-12 for x in f["features"])  # beta / number of matched features
    assert tuple(f["features"][0]["w"].shape) == (16, 8) and tuple(f["features"][1]["w"].shape) == (256, 64)
    a = AttentionDistillation(teacher, student, 2.0, 0.5, 0.2).fused_spec()
    assert abs(a["w_kd"] - 0.3) < 1e-12 and a["features"] == []
    u = UnifiedDistillation(teacher, student, 2.0, 0.4, 0.3, 0.2).fused_spec()
    assert abs(u["w_kd"] - 0.1) < 1e-12 and len(u["features"]) == 2
    assert all(not p.requires_grad for p in teacher.parameters())       # frozen (base.py:16-18)
    # equal architectures: tower activations match as well -> only the autograd path serves that pair
    twin = NCF(30, 20, 8, 3, 0.0, "NeuMF-end")
    same = FeatureDistillation(twin, student, 2.0, 0.5, 0.3)
    assert any(k.startswith("mlp_linear") for k in same.matched_keys())
    with pytest.raises(NotImplementedError):
        same.fused_spec()
