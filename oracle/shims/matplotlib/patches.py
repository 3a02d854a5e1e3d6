class Circle(object):
    pass


class Ellipse(object):
    pass
