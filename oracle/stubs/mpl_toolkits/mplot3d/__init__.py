class Axes3D(object):
    pass
