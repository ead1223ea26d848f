def save_checkpoint(*a, **k):
    raise NotImplementedError("refshim: checkpoints are not part of the executed path")


restore_checkpoint = save_checkpoint
