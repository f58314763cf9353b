from ncf_b200.distillation import BaseDistillation  # noqa: F401
