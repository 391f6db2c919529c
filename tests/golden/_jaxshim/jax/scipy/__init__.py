class signal:
    pass
