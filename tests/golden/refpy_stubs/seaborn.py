"""no-op stand-in"""


def histplot(*a, **k):
    return None
