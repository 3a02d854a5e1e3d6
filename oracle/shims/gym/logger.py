def set_level(level):
    return None
