"""Pins the CPU oracle (oracle/) against outputs of the reference itself (tests/golden/, produced
by oracle/make_golden.py from the unmodified reference) and against published known answers."""
import numpy as np
import pytest

from oracle import ncf_numpy as onp
from oracle import philox as oph
from tests.util import assert_close, assert_close_adam, group, load_golden

TRAIN_CASES = ["train_gmf_f8", "train_mlp_f8_l3", "train_neumf_f8_l3", "train_neumf_f32_l2",
               "train_neumf_f6_l2", "train_neumf_f5_l1", "train_neumf_f64_l3", "train_neumf_f8_l3_sgd",
               # shapes of the opt-in GPU sweep (tests/test_gpu_umma.py)
               "train_neumf_f32_l1", "train_mlp_f32_l3", "train_neumf_f64_l1", "train_neumf_f64_l2"]


def _batch(z, meta, t):
    n = meta["B"] - meta["short_last"] if (meta.get("short_last") and t == meta["T"] - 1) else meta["B"]
    return z["user"][t, :n], z["item"][t, :n], z["label"][t, :n]


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_forward_loss_grads_match_reference(name):
    z, meta = load_golden(name)
    params = group(z, "init")
    u, i, y = _batch(z, meta, 0)
    logits = onp.forward(params, u, i, meta["model_type"])
    assert_close(logits, z["logits0"], "logits")
    loss, dl = onp.loss_and_dlogit(logits, y)
    assert abs(float(loss) - z["loss"][0]) <= 1e-6 * abs(z["loss"][0])
    g = onp.backward(params, u, i, meta["model_type"], dl)
    ref_g = group(z, "grad0")
    assert set(g) == set(ref_g)  # same parameters receive a gradient as under autograd
    for k in ref_g:
        assert_close(g[k], ref_g[k], f"grad {k}")


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_optimizer_steps_match_reference(name):
    z, meta = load_golden(name)
    params = group(z, "init")
    opt = onp.DenseAdam(lr=meta["lr"]) if meta["optimizer"] == "adam" else None
    for t in range(meta["T"]):
        u, i, y = _batch(z, meta, t)
        logits = onp.forward(params, u, i, meta["model_type"])
        loss, dl = onp.loss_and_dlogit(logits, y)
        assert abs(float(loss) - z["loss"][t]) <= 2e-6 * abs(z["loss"][t]), f"loss at step {t}"
        g = onp.backward(params, u, i, meta["model_type"], dl)
        if opt is not None:
            opt.step(params, g)
        else:
            onp.sgd_step(params, g, meta["lr"])
        if t == 0:
            # Adam: an element whose gradient is at the eps = 1e-8 scale amplifies fp32 summation noise
            # (train_neumf_f32_l1 has one with g = 1.9e-8 in a row of 7e-6 gradients) - tests/util.py
            for k, ref in group(z, "after1").items():
                (assert_close_adam if opt is not None else assert_close)(params[k], ref, f"after step 1: {k}")
    check = assert_close_adam if opt is not None else assert_close
    for k, ref in group(z, "final").items():
        check(params[k], ref, f"final {k}")


def test_kd_response_matches_reference():
    z, meta = load_golden("kd_response")
    teacher, params = group(z, "teacher"), group(z, "init")
    opt = onp.DenseAdam(lr=meta["lr"])
    for t in range(meta["T"]):
        u, i, y = z["user"][t], z["item"][t], z["label"][t]
        tl = onp.forward(teacher, u, i, "NeuMF-end")
        sl = onp.forward(params, u, i, "NeuMF-end")
        if t == 0:
            assert_close(tl, z["teacher_logits0"], "teacher logits")
            assert_close(sl, z["student_logits0"], "student logits")
        loss, dl = onp.loss_and_dlogit(sl, y, tl, meta["alpha"])
        assert abs(float(loss) - z["loss"][t]) <= 2e-6 * abs(z["loss"][t])
        g = onp.backward(params, u, i, "NeuMF-end", dl)
        if t == 0:
            for k, ref in group(z, "grad0").items():
                assert_close(g[k], ref, f"KD grad {k}")
        opt.step(params, g)
    for k, ref in group(z, "final").items():
        assert_close_adam(params[k], ref, f"KD final {k}")


def test_pretrain_sgd_matches_reference():
    z, meta = load_golden("neumf_pre_sgd")
    params = group(z, "init")
    for t in range(meta["T"]):
        u, i, y = z["user"][t], z["item"][t], z["label"][t]
        logits = onp.forward(params, u, i, "NeuMF-pre")
        loss, dl = onp.loss_and_dlogit(logits, y)
        assert abs(float(loss) - z["loss"][t]) <= 2e-6 * abs(z["loss"][t])
        onp.sgd_step(params, onp.backward(params, u, i, "NeuMF-pre", dl), meta["lr"])
    for k, ref in group(z, "final").items():
        assert_close(params[k], ref, f"final {k}")


@pytest.mark.parametrize("name", ["metrics_neumf_f8_l3", "metrics_gmf_f8"])
def test_metrics_match_reference(name):
    z, meta = load_golden(name)
    # (1) the ranking rule on the reference's own scores reproduces its HR / NDCG lists exactly
    HR, NDCG = onp.metrics_from_scores(z["scores"], z["cands"], meta["k"])
    assert HR == z["HR"].tolist()
    assert NDCG == z["NDCG"].tolist()  # float64 1/log2(index+2), bit for bit
    # (2) scores recomputed by the oracle forward agree to tolerance
    (_, _), scores = onp.metrics(group(z, "init"), meta["model_type"], z["users"], z["cands"], meta["k"])
    assert_close(scores, z["scores"], "scores")


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = oph.philox4x32_10(*[np.uint32(x) for x in c], *[np.uint32(x) for x in k])
        assert tuple(int(x) for x in got) == want


def test_shuffle_perm_is_a_bijection():
    for S in (1, 2, 3, 7, 64, 1000, 4097):
        p = oph.shuffle_perm(S, seed=5, epoch=3, q=np.arange(S))
        assert sorted(p.tolist()) == list(range(S))
    a = oph.shuffle_perm(1000, 5, 3, np.arange(1000))
    b = oph.shuffle_perm(1000, 5, 4, np.arange(1000))
    assert (a != b).mean() > 0.9  # a new epoch reshuffles


def test_sampler_semantics_match_reference_sampler():
    """Same semantics as NCFData.ng_sample (datasets.py:53-69): never an observed pair, uniform
    over the user's non-interacted items, positives-then-negatives order, labels 1..1,0..0."""
    z, meta = load_golden("sampler_stats")
    pairs, ref_hist = z["pairs"], z["hist"]
    U, I, num_ng, reps = meta["U"], meta["I"], meta["num_ng"], meta["reps"]
    pu, pi = pairs[:, 0], pairs[:, 1]
    rowptr, col = oph.csr_build(pu, pi, U)
    hist = np.zeros((U, I), dtype=np.int64)
    for e in range(reps):
        neg = oph.sample_neg(rowptr, col, pu, num_ng, I, seed=99, epoch=e)
        np.add.at(hist, (np.repeat(pu, num_ng), neg), 1)
    observed = np.zeros((U, I), dtype=bool)
    observed[pu, pi] = True
    assert hist[observed].sum() == 0 and ref_hist[observed].sum() == 0
    assert hist.sum() == ref_hist.sum()
    # per-user draws are uniform over the free items: compare both samplers with the expectation
    for u in range(U):
        free = ~observed[u]
        n = hist[u].sum()
        if n == 0:
            continue
        exp = n / free.sum()
        for h in (hist[u][free], ref_hist[u][free]):
            chi2 = ((h - exp) ** 2 / exp).sum()
            assert chi2 < 3.0 * free.sum() + 20, (u, chi2)
    # stream order and labels (datasets.py:65-69)
    feats, labels = z["last_features"], z["last_labels"]
    P = pairs.shape[0]
    assert (labels[:P] == 1).all() and (labels[P:] == 0).all()
    assert (feats[:P] == pairs).all()
    assert (feats[P:, 0] == np.repeat(pu, num_ng)).all()
    su, si, sl = oph.shuffle_epoch(pu, pi, neg, num_ng, 99, 0, 0, P * (1 + num_ng))
    ours = sorted(zip(su.tolist(), si.tolist(), sl.tolist()))
    want = sorted(zip(np.concatenate([pu, np.repeat(pu, num_ng)]).tolist(),
                      np.concatenate([pi, neg]).tolist(), [1.0] * P + [0.0] * (P * num_ng)))
    assert ours == want  # the shuffled stream is a permutation of positives-then-negatives
