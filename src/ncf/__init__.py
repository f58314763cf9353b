"""Import shim over ncf_b200 (see src/__init__.py)."""
