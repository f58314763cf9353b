#!/usr/bin/env python3
"""Train the teacher NeuMF on the fused B200 path — CLI, epoch line and checkpoint name of the
reference's scripts/train_teacher.py (flags :113-125, `teacher_{model}_best.pth` :98).  The
reference sizes the tables from the YAML (`config.user_num/item_num`, :138-139) — kept, unless a
synthetic shape is named; TensorBoard scalars are written when tensorboardX is importable and a
per-epoch history JSON is written next to the logs either way.

    python scripts/train_teacher.py --model NeuMF-end --epochs 20
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.append(os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from ncf_b200.config import config
from ncf_b200.models import NCF
from ncf_b200.train_loop import fit, load_dataset


def _writer(name):
    try:
        from tensorboardX import SummaryWriter
        return SummaryWriter(log_dir=str(config.log_dir / name))
    except Exception:
        return None


def train_teacher(model_type, user_num, item_num, device, args):
    train, test_users, test_cands, data_users, data_items, _ = load_dataset(device, args.synthetic)
    if args.synthetic:
        user_num, item_num = data_users, data_items
    model = NCF(user_num, item_num, args.factor_num, args.num_layers, args.dropout, model_type)
    if model_type == "NeuMF-pre":
        gmf_path = config.model_dir / f"GMF_{args.factor_num}f_best.pth"
        mlp_path = config.model_dir / f"MLP_{args.num_layers}l_{args.factor_num}f_best.pth"
        if not (gmf_path.exists() and mlp_path.exists()):
            raise FileNotFoundError("Pretrained GMF or MLP weights not found")
        model.load_pretrain_weights(torch.load(gmf_path, map_location="cpu"),
                                    torch.load(mlp_path, map_location="cpu"))
        print("Loaded pretrained GMF and MLP weights for NeuMF-pre")
    model.to(device)
    config.ensure_dirs()
    run = f"teacher_{model_type}_{time.strftime('%Y%m%d_%H%M%S')}"
    writer = _writer(run)

    def on_epoch(epoch, loss, hr, ndcg, elapsed):
        if writer is not None:
            writer.add_scalar("Loss/Train", loss, epoch)
            writer.add_scalar(f"HR@{args.top_k}", hr, epoch)
            writer.add_scalar(f"NDCG@{args.top_k}", ndcg, epoch)
        print(f"Epoch {epoch+1:03d}: Loss={loss:.4f}, HR={hr:.3f}, NDCG={ndcg:.3f}, "
              f"Time={time.strftime('%H:%M:%S', time.gmtime(elapsed))}")

    def on_best(m):
        if args.out:
            path = config.model_dir / f"teacher_{model_type}_best.pth"
            torch.save(m.state_dict(), path)
            print(f"Saved best model to {path}")

    res = fit(model, train, test_users, test_cands, epochs=args.epochs, batch_size=args.batch_size,
              lr=args.lr, num_ng=args.num_ng, top_k=args.top_k, optimizer="adam", seed=args.seed,
              on_epoch=on_epoch, on_best=on_best)
    (config.log_dir / f"{run}_history.json").write_text(json.dumps(res.history, indent=1))
    if writer is not None:
        writer.close()
    return res.best_loss, res.best_hr, res.best_ndcg, res.best_epoch


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--lr", type=float, default=config.lr)
    p.add_argument("--dropout", type=float, default=config.dropout)
    p.add_argument("--batch_size", type=int, default=config.batch_size)
    p.add_argument("--epochs", type=int, default=config.epochs)
    p.add_argument("--top_k", type=int, default=config.top_k)
    p.add_argument("--factor_num", type=int, default=config.factor_num)
    p.add_argument("--num_layers", type=int, default=config.num_layers)
    p.add_argument("--num_ng", type=int, default=config.num_ng)
    p.add_argument("--test_num_ng", type=int, default=config.test_num_ng)
    p.add_argument("--out", action="store_true", default=True)
    p.add_argument("--gpu", type=str, default="0")
    p.add_argument("--model", type=str, default="NeuMF-end", choices=["NeuMF-end", "NeuMF-pre"])
    p.add_argument("--synthetic", type=str, default=None)
    p.add_argument("--seed", type=int, default=0)
    args = p.parse_args(argv)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)
    if not torch.cuda.is_available():
        raise SystemExit("ncf_b200 needs a CUDA device: there is no CPU fallback")
    print(f"Using GPU: {torch.cuda.get_device_name(0)}")
    best_loss, best_hr, best_ndcg, best_epoch = train_teacher(
        args.model, config.user_num, config.item_num, torch.device("cuda"), args)
    print(f"Best Epoch {best_epoch:03d}: Loss={best_loss:.4f}, HR={best_hr:.3f}, NDCG={best_ndcg:.3f}")
    return best_loss, best_hr, best_ndcg, best_epoch


if __name__ == "__main__":
    main()
