"""stub: imported but unused (state/ensemble.py:11)"""
