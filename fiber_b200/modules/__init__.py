from .fiber_module import FIBERTransformerSS  # noqa: F401
