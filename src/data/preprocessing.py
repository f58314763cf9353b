from ncf_b200.preprocessing import LeaveOneOutPreprocessor  # noqa: F401  (reference src/data/preprocessing.py:10)
