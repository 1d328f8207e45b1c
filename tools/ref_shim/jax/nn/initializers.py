from typing import Any, Callable

Initializer = Callable[..., Any]
