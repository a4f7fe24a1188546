"""Stand-in: sampling.py:6 imports `sample` but never calls it."""


def sample(*a, **k):
    raise NotImplementedError
