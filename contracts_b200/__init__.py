"""contracts_b200 — B200-native batched simulator for the `contracts` social-dilemma environments."""
__version__ = "0.1.0"
