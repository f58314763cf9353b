"""GPU parity of the data side: on-disk formats parsed on the device and the
leave-one-out preprocessing, against fixtures of the reference itself (load_all_small / preprocess_small,
oracle/make_golden_r2.py) and the oracle's restatement (oracle/textio.py).  Integer work: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import philox as oph
from oracle import textio
from tests.util import load_golden

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def to_dev(b: bytes):
    a = np.frombuffer(b, dtype=np.uint8)
    return torch.from_numpy(a.copy()).to(dev()) if a.size else torch.empty(0, dtype=torch.uint8, device=dev())


def test_load_all_matches_reference(tmp_path, monkeypatch):
    """The files the reference's load_all() read (datasets.py:9-36) give the same arrays here."""
    from ncf_b200 import datasets
    z, meta = load_golden("load_all_small")
    train = datasets.parse_train_rating_device(to_dev(z["train_bytes"].tobytes()))
    users, cands = datasets.parse_test_negative_device(to_dev(z["neg_bytes"].tobytes()), meta["C"])
    assert np.array_equal(train.cpu().numpy(), z["train_data"])
    want = z["test_data"].reshape(-1, meta["C"], 2)
    assert np.array_equal(users.cpu().numpy(), want[:, 0, 0]) and np.array_equal(cands.cpu().numpy(), want[:, :, 1])
    # through the files and the reference-named entry point
    tr, te = tmp_path / "u.train.rating", tmp_path / "u.test.negative"
    tr.write_bytes(z["train_bytes"].tobytes())
    te.write_bytes(z["neg_bytes"].tobytes())
    monkeypatch.setattr(datasets.config, "train_rating", tr)
    monkeypatch.setattr(datasets.config, "test_negative", te)
    train_data, test_data, user_num, item_num, train_mat = datasets.load_all()
    assert np.array_equal(train_data, z["train_data"]) and np.array_equal(test_data, z["test_data"])
    assert (user_num, item_num) == (meta["user_num"], meta["item_num"])
    keys = z["train_mat_keys"]
    assert len(train_mat) == keys.shape[0] and all((int(u), int(i)) in train_mat for u, i in keys[:50])
    host = datasets.load_all(host=True)
    assert np.array_equal(host[0], train_data) and np.array_equal(host[1], test_data)


def test_text_parser_edge_cases():
    from ncf_b200 import ops
    from ncf_b200._lib import NcfError
    from ncf_b200 import datasets
    # empty file, no trailing newline, CRLF, blank lines, extra columns, a negative number
    vals, status = ops.text_parse_ints(to_dev(b""), 2)
    assert vals.shape == (0, 2) and status == 0
    vals, status = ops.text_parse_ints(to_dev(b"1\t2\n3\t4"), 2)
    assert vals.tolist() == [[1, 2], [3, 4]] and status == 0
    vals, status = ops.text_parse_ints(to_dev(b"\n\n10\t20\t5\t881250949\r\n\r\n30\t40\t3\t7\n\n"), 2)
    assert vals.tolist() == [[10, 20], [30, 40]] and status == 0
    vals, status = ops.text_parse_ints(to_dev(b"7\t-3\n"), 2)
    assert vals.tolist() == [[7, -3]]
    vals, status = ops.text_parse_ints(to_dev(b"1\t2\t3\n4\n"), 3, exact=True)
    assert vals.tolist() == [[1, 2, 3], [4, -1, -1]] and status == 1
    vals, status = ops.text_parse_ints(to_dev(b"1 2 3 4\n"), 3, exact=True)
    assert status == 2
    with pytest.raises(NcfError):      # a user with 98 negatives: refused (the reference would misalign every later user)
        datasets.parse_test_negative_device(to_dev(b"(0,5)\t" + b"\t".join(str(k).encode() for k in range(6, 104)) + b"\n"), 100)
    # a large file against numpy: 300k lines, line lengths vary, chunk boundaries everywhere
    rng = np.random.default_rng(0)
    a = rng.integers(0, 10 ** rng.integers(1, 9, size=300_000), dtype=np.int64)
    b = rng.integers(0, 3000, size=300_000, dtype=np.int64)
    text = "".join(f"{x}\t{y}\n" for x, y in zip(a.tolist(), b.tolist())).encode()
    vals, status = ops.text_parse_ints(to_dev(text), 2)
    assert status == 0 and np.array_equal(vals.cpu().numpy(), np.stack([a, b], 1))


def test_leave_one_out_matches_reference_preprocessor(tmp_path):
    from ncf_b200 import ops
    from ncf_b200.preprocessing import LeaveOneOutPreprocessor
    z, meta = load_golden("preprocess_small")
    raw = tmp_path / "u.data"
    raw.write_bytes(z["raw_bytes"].tobytes())
    pre = LeaveOneOutPreprocessor(raw_path=raw, processed_dir=tmp_path / "processed", num_negatives=99, seed=11)
    df = pre.load_and_prepare_data()
    train, test = pre.temporal_split(df)
    assert np.array_equal(train.cpu().numpy(), z["train_data"]) and np.array_equal(test.cpu().numpy(), z["test_data"])
    # negatives: bit-exact against the oracle's restatement of the draw rule, and the reference's contract
    num_items = meta["num_items"]
    negs = pre.generate_test_negatives(train, test, num_items).cpu().numpy()
    allp = np.concatenate([z["train_data"], z["test_data"]])
    rowptr, col = oph.csr_build(allp[:, 0], allp[:, 1], meta["num_users"])
    want, cnt = textio.eval_negatives(rowptr, col, z["test_data"][:, 0], num_items, 99, seed=11)
    assert np.array_equal(negs, want) and (cnt == 99).all()
    assert (np.diff(negs, axis=1) > 0).all() and negs.min() >= 0 and negs.max() < num_items
    for r, u in enumerate(z["test_data"][:, 0]):
        assert not np.isin(negs[r], col[rowptr[u]:rowptr[u + 1]]).any()
    # the whole pipeline: same files as the reference wrote, and they load back
    info = pre.run()
    assert info["train_interactions"] == len(z["train_data"]) and info["test_interactions"] == len(z["test_data"])
    assert (tmp_path / "processed" / "u.train.rating").read_bytes() == z["train_file"].tobytes()
    assert (tmp_path / "processed" / "u.test.rating").read_bytes() == z["test_rating_file"].tobytes()
    neg_text = (tmp_path / "processed" / "u.test.negative").read_bytes()
    assert not neg_text.endswith(b"\n")
    back = textio.parse_test_negative(neg_text).reshape(-1, 100, 2)
    assert np.array_equal(back[:, 0], z["test_data"]) and np.array_equal(back[:, 1:, 1], want)


def test_split_and_negatives_randomised_against_oracle():
    """Ties in the timestamps (file order decides), a single-rating user, an absent user, a user with 9 000
    ratings (one CTA sorts the row), 200k ratings in all."""
    from ncf_b200 import ops
    rng = np.random.default_rng(1)
    U, I, n = 3000, 12000, 200_000
    user = rng.integers(0, U, n)
    user[user == 17] = 18                        # user 17 absent
    user[:9000] = 5                              # a very active user
    item = rng.integers(0, I, n)
    ts = rng.integers(1_000_000, 1_000_400, n)   # many ties
    user = np.concatenate([user, [17 + 2000]])   # exactly one rating for user 2017 (if none was drawn)
    keep = np.ones(user.shape[0], bool)
    keep[:-1] &= user[:-1] != 2017
    user, item, ts = user[keep], np.concatenate([item, [3]])[keep], np.concatenate([ts, [1_000_100]])[keep]
    train, test = ops.leave_one_out_split(*(torch.from_numpy(a).to(dev()) for a in (user, item, ts)), U)
    w_train, w_test = textio.temporal_split(user, item, ts)
    assert np.array_equal(train.cpu().numpy(), w_train) and np.array_equal(test.cpu().numpy(), w_test)
    assert 2017 not in w_test[:, 0] and 2017 in w_train[:, 0]
    allp = np.concatenate([w_train, w_test])
    rowptr, col = ops.csr_build(torch.from_numpy(allp[:, 0]).to(dev()), torch.from_numpy(allp[:, 1]).to(dev()), U)
    o_rowptr, o_col = oph.csr_build(allp[:, 0], allp[:, 1], U)
    sel = np.concatenate([np.arange(40), [np.nonzero(w_test[:, 0] == 5)[0][0]]])    # the oracle loops in Python: a sample
    tu = w_test[sel, 0]
    negs, cnt = ops.eval_negatives(rowptr, col, torch.from_numpy(tu).to(dev()), I, 99, seed=2**40 + 5)
    want, wcnt = textio.eval_negatives(o_rowptr, o_col, tu, I, 99, seed=2**40 + 5)
    assert np.array_equal(negs.cpu().numpy(), want) and np.array_equal(cnt.cpu().numpy(), wcnt)
    # a user who owns nearly every item runs out of draws: fewer than K negatives, -1 padded, like the reference's warning
    few_u = np.zeros(1, dtype=np.int64)
    rp, cl = ops.csr_build(torch.zeros(30, dtype=torch.int64, device=dev()), torch.arange(30, device=dev()), 1)
    negs, cnt = ops.eval_negatives(rp, cl, torch.from_numpy(few_u).to(dev()), 32, 5, seed=1)
    o_rp, o_cl = oph.csr_build(np.zeros(30, dtype=np.int64), np.arange(30), 1)
    want, wcnt = textio.eval_negatives(o_rp, o_cl, few_u, 32, 5, seed=1)
    assert np.array_equal(negs.cpu().numpy(), want) and int(cnt[0]) == int(wcnt[0]) <= 2
