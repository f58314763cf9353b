from ncf_b200.distillation import ResponseDistillation, SoftTargetDistillation  # noqa: F401
