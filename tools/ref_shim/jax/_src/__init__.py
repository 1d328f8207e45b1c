from . import core, dtypes  # noqa: F401
