"""no-op stand-in"""
PatchCollection = object
