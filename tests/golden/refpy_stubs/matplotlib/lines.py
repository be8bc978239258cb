"""no-op stand-in"""
