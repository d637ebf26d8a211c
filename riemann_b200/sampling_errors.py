"""Exceptions -- same names and behaviour as riemann/sampling_errors.py:11-28."""


class RiemannBaseError(Exception):
    def __init__(self, msg):
        self.msg = msg

    def __str__(self):
        return "{}: {}".format(self.__class__.__name__, self.msg)


class ParameterError(RiemannBaseError):
    """Faulty parameters (shapes, ranges, unsupported model/proposal pairs)."""
