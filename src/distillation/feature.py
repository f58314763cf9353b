from ncf_b200.distillation import FeatureDistillation  # noqa: F401
