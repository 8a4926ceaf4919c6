"""stub: plotting is out of scope (state/ensemble.py:9)"""
