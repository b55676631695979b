"""Version of the B200 re-implementation; tracks the reference release it mirrors (0.5.1)."""
__version__ = "0.5.1+b200.1"
