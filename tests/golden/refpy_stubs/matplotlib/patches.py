"""no-op stand-in"""
Circle = Wedge = Polygon = object
