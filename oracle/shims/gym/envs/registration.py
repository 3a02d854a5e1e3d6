def register(id=None, entry_point=None, **kwargs):
    import gym
    gym._REGISTRY[id] = entry_point
