"""Import shim over ncf_b200 (see src/__init__.py): the names reference src/distillation/__init__.py exports,
plus the UnifiedDistillation its scripts/train_student.py:19 imports."""
from ncf_b200.distillation import (AttentionDistillation, BaseDistillation, FeatureDistillation,  # noqa: F401
                                   ResponseDistillation, SoftTargetDistillation, UnifiedDistillation)
