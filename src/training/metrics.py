from ncf_b200.metrics import metrics  # noqa: F401  (reference src/training/metrics.py:4)
