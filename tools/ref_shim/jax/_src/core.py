class Shape(tuple):
    pass


def canonicalize_shape(s):
    return tuple(s)
