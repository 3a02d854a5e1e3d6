def ion():
    return None
