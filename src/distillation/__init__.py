from ncf_b200.distillation import BaseDistillation, ResponseDistillation  # noqa: F401
