"""gavisunk_b200 -- B200-native engine for GAVISUNK's SUNK match + inter-SUNK validation path."""
__version__ = "0.1.0"
