"""Import shim: lets code written against the reference (`from src.ncf.models import NCF`, ...) run on ncf_b200 unchanged."""
