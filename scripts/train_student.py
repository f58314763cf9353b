#!/usr/bin/env python3
"""Distil a teacher NeuMF into a smaller student on the fused B200 path — CLI, epoch line and
checkpoint names of the reference's scripts/train_student.py (flags :23-61, teacher =
(2*factor_num, num_layers+1) of the student :86, `student_{model}_best.pth` :179).

Only the response (teacher-score) strategy is on the hot path (BASELINE north_star); the
reference's feature / attention strategies are out of scope and `unified` does not exist in the
reference at HEAD — those choices exit with a clear message.  Decoupled
--teacher_factor_num / --teacher_num_layers default to the reference rule.

    python scripts/train_student.py --distillation response --epochs 20
"""
import argparse
import os
import sys
import time

import torch

sys.path.append(os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from ncf_b200.config import config
from ncf_b200.models import NCF
from ncf_b200.train_loop import fit, load_dataset


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--lr", type=float, default=config.lr)
    p.add_argument("--dropout", type=float, default=config.dropout)
    p.add_argument("--batch_size", type=int, default=config.batch_size)
    p.add_argument("--epochs", type=int, default=config.epochs)
    p.add_argument("--top_k", type=int, default=config.top_k)
    p.add_argument("--factor_num", type=int, default=config.factor_num // 2)
    p.add_argument("--num_layers", type=int, default=config.num_layers - 1)
    p.add_argument("--num_ng", type=int, default=config.num_ng)
    p.add_argument("--test_num_ng", type=int, default=config.test_num_ng)
    p.add_argument("--out", action="store_true", default=True)
    p.add_argument("--gpu", type=str, default="0")
    p.add_argument("--teacher_model", type=str, default=config.model_type,
                   choices=["GMF", "MLP", "NeuMF-end", "NeuMF-pre"])
    p.add_argument("--student_model", type=str, default="NeuMF-end", choices=["GMF", "MLP", "NeuMF-end"])
    p.add_argument("--temperature", type=float, default=config.temperature)
    p.add_argument("--alpha", type=float, default=config.alpha)
    p.add_argument("--beta", type=float, default=0.3)
    p.add_argument("--gamma", type=float, default=0.2)
    p.add_argument("--distillation", type=str, default="response",
                   choices=["response", "feature", "attention", "unified"])
    p.add_argument("--teacher_factor_num", type=int, default=None)
    p.add_argument("--teacher_num_layers", type=int, default=None)
    p.add_argument("--synthetic", type=str, default=None)
    p.add_argument("--seed", type=int, default=0)
    args = p.parse_args(argv)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", args.gpu)
    if not torch.cuda.is_available():
        raise SystemExit("ncf_b200 needs a CUDA device: there is no CPU fallback")
    device = torch.device("cuda")
    print(" Using GPU:", torch.cuda.get_device_name(0))

    train, test_users, test_cands, user_num, item_num, _ = load_dataset(device, args.synthetic)
    teacher_path = config.model_dir / f"teacher_{args.teacher_model}_best.pth"
    assert os.path.exists(teacher_path), f"Lack of teacher model: {teacher_path}"
    tf = args.teacher_factor_num or args.factor_num * 2
    tl = args.teacher_num_layers or args.num_layers + 1
    teacher = NCF(user_num, item_num, tf, tl, args.dropout, args.teacher_model)
    teacher.load_state_dict(torch.load(teacher_path, map_location="cpu"))
    teacher.to(device).eval()
    student = NCF(user_num, item_num, args.factor_num, args.num_layers, args.dropout, args.student_model).to(device)

    def on_epoch(epoch, loss, hr, ndcg, elapsed):
        print(f"{epoch:03d} - Loss: {loss:.6f}, HR: {hr:.3f}, NDCG: {ndcg:.3f}, "
              f"Time: {time.strftime('%H:%M:%S', time.gmtime(elapsed))}")

    def on_best(m):
        if args.out:
            config.ensure_dirs()
            path = config.model_dir / f"student_{args.student_model}_best.pth"
            torch.save(m.state_dict(), path)
            print(f"Saved best model to {path}")

    # reference scripts/train_student.py:96-127 (its fourth branch imports a class its own tree does not have)
    from ncf_b200.distillation import (AttentionDistillation, FeatureDistillation, ResponseDistillation,
                                       UnifiedDistillation)
    if args.distillation == "response":
        distillation = ResponseDistillation(teacher, student, temperature=args.temperature, alpha=args.alpha)
    elif args.distillation == "feature":
        distillation = FeatureDistillation(teacher, student, temperature=args.temperature, alpha=args.alpha,
                                           beta=args.beta)
    elif args.distillation == "attention":
        distillation = AttentionDistillation(teacher, student, temperature=args.temperature, alpha=args.alpha,
                                             gamma=args.gamma)
    else:
        distillation = UnifiedDistillation(teacher, student, temperature=args.temperature, alpha=args.alpha,
                                           beta=args.beta, gamma=args.gamma)
    distillation.to(device)
    res = fit(student, train, test_users, test_cands, epochs=args.epochs, batch_size=args.batch_size,
              lr=args.lr, num_ng=args.num_ng, top_k=args.top_k, optimizer="adam", distillation=distillation,
              seed=args.seed, on_epoch=on_epoch, on_best=on_best)
    print(f"End. Best epoch {res.best_epoch:03d}: HR = {res.best_hr:.3f}, NDCG = {res.best_ndcg:.3f}")
    return res


if __name__ == "__main__":
    main()
