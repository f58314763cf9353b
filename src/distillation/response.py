from ncf_b200.distillation import ResponseDistillation  # noqa: F401
