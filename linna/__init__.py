"""Drop-in import name.  Reference pickles (``X_transform.pkl``, ``model_args.pkl`` ...) name
their classes by module path ``linna.util`` / ``linna.nn``, so the B200 package is
also importable as ``linna``.  Everything here re-exports ``linna_b200``."""
__author__ = """linna_b200"""
__version__ = "0.1.0"
