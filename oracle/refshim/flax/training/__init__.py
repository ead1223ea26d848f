from . import checkpoints, train_state  # noqa: F401
