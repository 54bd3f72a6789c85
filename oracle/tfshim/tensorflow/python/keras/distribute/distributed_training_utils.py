def is_tpu_strategy(strategy):
    return False
