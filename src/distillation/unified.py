from ncf_b200.distillation import UnifiedDistillation  # noqa: F401
