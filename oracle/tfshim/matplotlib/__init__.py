"""Empty stand-in so that the reference's graph_func.py (plot helpers, unused on the hot path) can be imported."""
