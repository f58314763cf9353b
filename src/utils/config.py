from ncf_b200.config import Config, config  # noqa: F401  (reference src/utils/config.py:65)
