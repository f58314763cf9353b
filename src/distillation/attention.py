from ncf_b200.distillation import AttentionDistillation  # noqa: F401
