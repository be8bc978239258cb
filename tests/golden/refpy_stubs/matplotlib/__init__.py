"""no-op stand-in: the reference scripts import matplotlib / seaborn for figures that the hot path never writes"""
from . import pyplot, lines, ticker, patches, collections  # noqa: F401
