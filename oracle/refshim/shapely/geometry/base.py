class BaseGeometry(object):
    pass
