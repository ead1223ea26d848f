import dataclasses
from typing import Any, Callable


@dataclasses.dataclass
class TrainState:
    step: int = 0
    apply_fn: Callable = None
    params: Any = None
    tx: Any = None
    opt_state: Any = None

    def replace(self, **kw):
        return dataclasses.replace(self, **kw)
