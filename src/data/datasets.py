from ncf_b200.datasets import NCFData, load_all  # noqa: F401  (reference src/data/datasets.py:9,39)
