from ncf_b200.models import NCF  # noqa: F401  (reference src/ncf/models.py:4)
