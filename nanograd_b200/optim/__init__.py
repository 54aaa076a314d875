from nanograd_b200.optim.optimizer import SGD, Adam, AdamW, Optimizer  # noqa: F401
