#!/usr/bin/env python3
"""Train NeuMF on the fused B200 path — same CLI, stdout contract and checkpoint names as the
reference's scripts/train_neumf.py (flags :170-196, epoch line :131, RESULTS block :150-157,
`NeuMF_{end|pre}_{L}l_{f}f_best.pth` :139-141).

    python scripts/train_neumf.py --model NeuMF-end --num_layers 3
    python scripts/train_neumf.py --synthetic ml1m --epochs 2      # no data files needed
"""
import argparse
import os
import sys

import torch

sys.path.append(os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from ncf_b200.config import config
from ncf_b200.models import NCF
from ncf_b200.train_loop import count_parameters, fit, load_dataset


def find_pretrained_model(model_type, num_layers, factor_num):
    name = f"GMF_{factor_num}f_best.pth" if model_type == "GMF" else f"MLP_{num_layers}l_{factor_num}f_best.pth"
    path = config.model_dir / name
    return path if path.exists() else None


def train_neumf(args, device):
    print(f"\nTraining {args.model} with {args.num_layers} layers...")
    print(f"Pretraining: {'Yes' if args.pretraining else 'No'}")
    train, test_users, test_cands, user_num, item_num, _ = load_dataset(device, args.synthetic)
    print(f"Dataset: {user_num} users, {item_num} items")
    model = NCF(user_num, item_num, args.factor_num, args.num_layers, args.dropout, args.model)
    if args.pretraining:
        gmf_path = find_pretrained_model("GMF", None, args.factor_num)
        mlp_path = find_pretrained_model("MLP", args.num_layers, args.factor_num)
        if gmf_path and mlp_path:
            print("Loading pretrained weights...")
            model.load_pretrain_weights(torch.load(gmf_path, map_location="cpu"),
                                        torch.load(mlp_path, map_location="cpu"))
            print("Pretrained weights loaded successfully")
        else:
            print("Warning: Pretrained weights not found!")
            print(f"GMF path: {gmf_path}")
            print(f"MLP path: {mlp_path}")
            print("Training without pretraining...")
            args.pretraining = False
    model.to(device)
    model.tower_math = args.tower_math
    param_count = count_parameters(model)
    print(f"Model parameters: {param_count:,}")
    # SGD(lr*10) for pretrained models as in the reference (train_neumf.py:87-90), Adam otherwise
    optimizer, lr = ("sgd", args.lr * 10) if args.pretraining else ("adam", args.lr)
    suffix = "pre" if args.pretraining else "end"
    model_filename = f"NeuMF_{suffix}_{args.num_layers}l_{args.factor_num}f_best.pth"

    def on_epoch(epoch, loss, hr, ndcg, elapsed):
        print(f"Epoch {epoch+1:03d}: Loss={loss:.4f}, HR={hr:.3f}, NDCG={ndcg:.3f}, Time={elapsed:.1f}s")

    def on_best(m):
        if args.save:
            config.ensure_dirs()
            torch.save(m.state_dict(), config.model_dir / model_filename)
            print(f"Saved best model: {model_filename}")

    print(f"Training for {args.epochs} epochs...")
    res = fit(model, train, test_users, test_cands, epochs=args.epochs, batch_size=args.batch_size,
              lr=lr, num_ng=args.num_ng, top_k=args.top_k, optimizer=optimizer, seed=args.seed,
              on_epoch=on_epoch, on_best=on_best)
    best_epoch = res.best_epoch + 1 if res.best_hr > 0 else 0
    print("\nTraining completed!")
    print(f"Best Result: Epoch {best_epoch:03d}: HR={res.best_hr:.3f}, NDCG={res.best_ndcg:.3f}")
    print("\n--- RESULTS ---")
    print(f"Model: {args.model}")
    print(f"Layers: {args.num_layers}")
    print(f"Pretraining: {args.pretraining}")
    print(f"HR@{args.top_k}: {res.best_hr}")
    print(f"NDCG@{args.top_k}: {res.best_ndcg}")
    print(f"Parameters: {param_count}")
    print("--- END RESULTS ---")
    return {"best_hr": res.best_hr, "best_ndcg": res.best_ndcg, "best_epoch": best_epoch,
            "parameters": param_count, "num_layers": args.num_layers, "pretraining": args.pretraining,
            "model_type": args.model}


def main(argv=None):
    p = argparse.ArgumentParser(description="Train NeuMF model")
    p.add_argument("--model", type=str, default="NeuMF-end", choices=["NeuMF-end", "NeuMF-pre"])
    p.add_argument("--epochs", type=int, default=config.epochs)
    p.add_argument("--factor_num", type=int, default=config.factor_num)
    p.add_argument("--num_layers", type=int, default=config.num_layers)
    p.add_argument("--pretraining", action="store_true")
    p.add_argument("--lr", type=float, default=config.lr)
    p.add_argument("--batch_size", type=int, default=config.batch_size)
    p.add_argument("--dropout", type=float, default=config.dropout)
    p.add_argument("--num_ng", type=int, default=config.num_ng)
    p.add_argument("--test_num_ng", type=int, default=config.test_num_ng)
    p.add_argument("--top_k", type=int, default=config.top_k)
    p.add_argument("--save", action="store_true", default=True)
    p.add_argument("--gpu", type=str, default="0")
    # additions (not in the reference)
    p.add_argument("--synthetic", type=str, default=None, help="synthetic shape (ml100k|ml1m|ml20m) instead of data files")
    p.add_argument("--seed", type=int, default=0, help="sampler / shuffle seed")
    p.add_argument("--tower_math", choices=["fp32", "tf32"], default="fp32")
    args = p.parse_args(argv)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)
    if not torch.cuda.is_available():
        raise SystemExit("ncf_b200 needs a CUDA device: there is no CPU fallback")
    print(f"Using GPU: {torch.cuda.get_device_name(0)}")
    result = train_neumf(args, torch.device("cuda"))
    print("\nFinal Results:")
    print(f"Model: {result['model_type']}")
    print(f"Layers: {result['num_layers']}")
    print(f"Pretraining: {result['pretraining']}")
    print(f"HR@{args.top_k}: {result['best_hr']:.4f}")
    print(f"NDCG@{args.top_k}: {result['best_ndcg']:.4f}")
    print(f"Parameters: {result['parameters']:,}")
    return result


if __name__ == "__main__":
    main()
