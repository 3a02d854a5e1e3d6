class Box(object):
    def __init__(self, low=None, high=None, shape=None, dtype=None):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class Dict(object):
    def __init__(self, spaces=None):
        self.spaces = spaces
