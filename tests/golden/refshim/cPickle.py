"""stub: Python-2 module imported but unused (observation/observation.py:6)"""
