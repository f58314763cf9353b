"""CPU: the oracle's restatements of the distillation objectives, the K sweep, the on-disk formats, the
leave-one-out split and the training loop against fixtures produced by the UNMODIFIED reference
(oracle/make_golden_r2.py -> tests/golden/*.npz)."""
import numpy as np
import pytest

from oracle import ncf_numpy as onp
from oracle import philox as oph
from oracle import textio
from tests.util import assert_close, assert_close_adam, group, load_golden

KD_CASES = ["kd_response_cfg3", "kd_soft_target", "kd_feature", "kd_feature_same", "kd_attention"]


def kd_objective(meta, z, student, teacher, u, i, y):
    """(loss, gradients of the student) of the case's distillation objective, from the oracle."""
    kind, a, T = meta["kind"], meta["alpha"], meta["temperature"]
    t_logits = onp.forward(teacher, u, i, "NeuMF-end")
    s_logits = onp.forward(student, u, i, "NeuMF-end")
    dfeat, extra = None, 0.0
    if kind == "response":
        loss, dl = onp.kd_loss_and_dlogit(s_logits, y, t_logits, a, 1 - a, T, 0)
    elif kind == "soft":
        loss, dl = onp.kd_loss_and_dlogit(s_logits, y, t_logits, a, 1 - a, T, 1)
    elif kind == "feature":
        b = meta["beta"]
        loss, dl = onp.kd_loss_and_dlogit(s_logits, y, t_logits, a, max(0, 1 - a - b), T, 1)
        adapters = {}
        for key in ("gmf_features", "mlp_input"):
            if f"adapter/{key}.weight" in z.files:
                adapters[key] = (z[f"adapter/{key}.weight"], z[f"adapter/{key}.bias"])
        extra, dfeat = onp.feature_matching(student, teacher, u, i, adapters, b)
    else:
        g = meta["gamma"]
        loss, dl = onp.kd_loss_and_dlogit(s_logits, y, t_logits, a, 1 - a - g, T, 1)
        extra = g * onp.attention_transfer(student, teacher, u, i)
    grads = onp.backward(student, u, i, "NeuMF-end", dl, dfeat)
    return float(loss) + extra, grads, t_logits, s_logits


@pytest.mark.parametrize("name", KD_CASES)
def test_distillation_objectives_match_reference(name):
    z, meta = load_golden(name)
    teacher, student = group(z, "teacher"), group(z, "init")
    opt = onp.DenseAdam(lr=meta["lr"])
    for t in range(meta["T"]):
        u, i, y = z["user"][t], z["item"][t], z["label"][t]
        loss, grads, t_logits, s_logits = kd_objective(meta, z, student, teacher, u, i, y)
        assert abs(loss - z["loss"][t]) <= 2e-6 * abs(z["loss"][t]), (t, loss, z["loss"][t])
        if t == 0:
            assert_close(t_logits, z["teacher_logits0"], "teacher logits")
            assert_close(s_logits, z["student_logits0"], "student logits")
            for k, want in group(z, "grad0").items():
                assert_close(grads[k], want, f"grad {k}")
        opt.step(student, grads)
    for k, want in group(z, "final").items():
        assert_close_adam(student[k], want, f"final {k}")


def test_attention_transfer_is_numerically_nothing():
    """The reference's attention map is the constant 1/B (softmax over the batch of unit norms): the
    transfer term is ~1e-9 and carries no gradient."""
    z, meta = load_golden("kd_attention")
    v = onp.attention_transfer(group(z, "init"), group(z, "teacher"), z["user"][0], z["item"][0])
    assert abs(v) < 1e-6


def test_metrics_k_sweep_matches_reference():
    z, meta = load_golden("metrics_ksweep")
    got = onp.metrics_at_k(group(z, "init"), "NeuMF-end", z["users"], z["cands"], range(1, 11))
    assert np.allclose([got[k][0] for k in range(1, 11)], z["hr_at_k"], atol=0, rtol=0)
    assert np.allclose([got[k][1] for k in range(1, 11)], z["ndcg_at_k"], rtol=1e-12)


def test_file_formats_match_reference_load_all():
    z, meta = load_golden("load_all_small")
    train = textio.parse_train_rating(z["train_bytes"].tobytes())
    test = textio.parse_test_negative(z["neg_bytes"].tobytes())
    assert np.array_equal(train, z["train_data"]) and np.array_equal(test, z["test_data"])
    assert train[:, 0].max() + 1 == meta["user_num"] and train[:, 1].max() + 1 == meta["item_num"]
    keys = np.unique(train, axis=0)
    assert np.array_equal(keys, z["train_mat_keys"])          # the dok_matrix's key set


def test_leave_one_out_split_matches_reference_preprocessor():
    z, meta = load_golden("preprocess_small")
    raw = np.array([ln.split("\t") for ln in z["raw_bytes"].tobytes().decode().splitlines()], dtype=np.int64)
    train, test = textio.temporal_split(raw[:, 0], raw[:, 1], raw[:, 3])
    assert np.array_equal(train, z["train_data"]) and np.array_equal(test, z["test_data"])
    # the files the reference wrote are exactly these rows, tab separated
    assert "".join(f"{u}\t{i}\n" for u, i in train).encode() == z["train_file"].tobytes()
    assert "".join(f"{u}\t{i}\n" for u, i in test).encode() == z["test_rating_file"].tobytes()
    # evaluation negatives: the reference's are RNG-dependent (numpy global MT19937); both must satisfy the
    # same contract: 99 distinct items, ascending, none of the user's items
    ni = meta["num_items"]
    allp = np.concatenate([train, test])
    rowptr, col = oph.csr_build(allp[:, 0], allp[:, 1], meta["num_users"])
    ours, cnt = textio.eval_negatives(rowptr, col, test[:, 0], ni, 99, seed=11)
    for negs in (ours, z["ref_negatives"]):
        assert negs.shape == (test.shape[0], 99)
        assert (np.diff(negs, axis=1) > 0).all() and negs.min() >= 0 and negs.max() < ni
        for r, u in enumerate(test[:, 0]):
            assert not np.isin(negs[r], col[rowptr[u]:rowptr[u + 1]]).any()
    assert (cnt == 99).all()
    # uniform over the free items: both samplers cover them alike (coarse chi-square-like bound)
    for negs in (ours, z["ref_negatives"]):
        h = np.bincount(negs.reshape(-1), minlength=ni)[1:]   # item id 0 does not exist in this data set
        assert h.std() / h.mean() < 0.35


def test_training_loop_matches_reference_first_epoch():
    """Epoch 0 of the quality fixture through the oracle: same batches (oracle/philox.py), same loss and
    HR@10 / NDCG@10 as the reference loop printed."""
    z, meta = load_golden("quality_ml100k")
    U, I, B, num_ng, seed = meta["U"], meta["I"], meta["B"], meta["num_ng"], meta["seed"]
    pu, pi = z["pos_user"].astype(np.int64), z["pos_item"].astype(np.int64)
    params = group(z, "init")
    rowptr, col = oph.csr_build(pu, pi, U)
    neg = oph.sample_neg(rowptr, col, pu, num_ng, I, seed, 0)
    S = pu.shape[0] * (1 + num_ng)
    su, si, sl = oph.shuffle_epoch(pu, pi, neg, num_ng, seed, 0, 0, S)
    opt = onp.DenseAdam(lr=meta["lr"])
    total, nb = 0.0, 0
    for q in range(0, S, B):
        u, i, y = su[q:q + B], si[q:q + B], sl[q:q + B]
        logits = onp.forward(params, u, i, "NeuMF-end")
        loss, dl = onp.loss_and_dlogit(logits, y)
        opt.step(params, onp.backward(params, u, i, "NeuMF-end", dl))
        total += float(loss)
        nb += 1
    cands = z["cands"].astype(np.int64)
    (HR, NDCG), _ = onp.metrics(params, "NeuMF-end", np.arange(U), cands, meta["top_k"])
    want = z["history"][0]
    assert abs(total / nb - want[0]) <= 1e-4 * want[0]
    assert abs(np.mean(HR) - want[1]) <= 0.005 and abs(np.mean(NDCG) - want[2]) <= 0.005
