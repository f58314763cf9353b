#!/usr/bin/env python3
"""Pretrain GMF or MLP on the fused B200 path — same CLI, stdout contract and checkpoint names
as the reference's scripts/pretrain.py (flags :113-137, RESULTS block :167-171,
`GMF_{f}f_best.pth` / `MLP_{L}l_{f}f_best.pth` :97-102).

    python scripts/pretrain.py --model GMF --epochs 20
"""
import argparse
import os
import sys

import torch

sys.path.append(os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from ncf_b200.config import config
from ncf_b200.models import NCF
from ncf_b200.train_loop import count_parameters, fit, load_dataset


def train_model(model_type, args, device):
    print(f"\nTraining {model_type} model...")
    train, test_users, test_cands, user_num, item_num, _ = load_dataset(device, args.synthetic)
    print("\nDataset Info:")
    print(f"Users: {user_num}, Items: {item_num}")
    print(f"Training interactions: {train.shape[0]}")
    model = NCF(user_num, item_num, args.factor_num, args.num_layers, args.dropout, model_type).to(device)
    param_count = count_parameters(model)
    print(f"Model: {model_type}")
    print(f"Parameters: {param_count:,}")
    print(f"Factor num: {args.factor_num}")
    print(f"Layers: {args.num_layers}")
    name = (f"MLP_{args.num_layers}l_{args.factor_num}f_best.pth" if model_type == "MLP"
            else f"{model_type}_{args.factor_num}f_best.pth")

    def on_epoch(epoch, loss, hr, ndcg, elapsed):
        print(f"Epoch {epoch+1:03d}: Loss={loss:.4f}, HR={hr:.3f}, NDCG={ndcg:.3f}, Time={elapsed:.1f}s")

    def on_best(m):
        if args.save:
            config.ensure_dirs()
            torch.save(m.state_dict(), config.model_dir / name)
            print(f"Saved best model to {config.model_dir / name}")

    print(f"Training for {args.epochs} epochs...")
    res = fit(model, train, test_users, test_cands, epochs=args.epochs, batch_size=args.batch_size,
              lr=args.lr, num_ng=args.num_ng, top_k=args.top_k, optimizer="adam", seed=args.seed,
              on_epoch=on_epoch, on_best=on_best)
    print("\nTraining completed!")
    print(f"Best HR@{args.top_k}: {res.best_hr:.4f} at epoch {res.best_epoch+1}")
    return res.best_hr, res.best_ndcg, param_count


def main(argv=None):
    p = argparse.ArgumentParser(description="Train GMF or MLP model")
    p.add_argument("--model", type=str, required=True, choices=["GMF", "MLP"])
    p.add_argument("--epochs", type=int, default=config.epochs)
    p.add_argument("--lr", type=float, default=config.lr)
    p.add_argument("--dropout", type=float, default=config.dropout)
    p.add_argument("--batch_size", type=int, default=config.batch_size)
    p.add_argument("--top_k", type=int, default=config.top_k)
    p.add_argument("--factor_num", type=int, default=config.factor_num)
    p.add_argument("--num_layers", type=int, default=config.num_layers)
    p.add_argument("--num_ng", type=int, default=config.num_ng)
    p.add_argument("--test_num_ng", type=int, default=config.test_num_ng)
    p.add_argument("--save", action="store_true", default=True)
    p.add_argument("--gpu", type=str, default="0")
    p.add_argument("--synthetic", type=str, default=None)
    p.add_argument("--seed", type=int, default=0)
    args = p.parse_args(argv)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)
    if not torch.cuda.is_available():
        raise SystemExit("ncf_b200 needs a CUDA device: there is no CPU fallback")
    print(f"Using GPU: {torch.cuda.get_device_name(0)}")
    best_hr, best_ndcg, param_count = train_model(args.model, args, torch.device("cuda"))
    print("\nFinal Results:")
    print(f"Model: {args.model}")
    print(f"HR@{args.top_k}: {best_hr:.4f}")
    print(f"NDCG@{args.top_k}: {best_ndcg:.4f}")
    print(f"Parameters: {param_count:,}")
    print("\n--- RESULTS ---")
    print(f"HR@{args.top_k}: {best_hr:.6f}")
    print(f"NDCG@{args.top_k}: {best_ndcg:.6f}")
    print(f"Parameters: {param_count}")
    print("--- END RESULTS ---")
    return best_hr, best_ndcg, param_count


if __name__ == "__main__":
    main()
