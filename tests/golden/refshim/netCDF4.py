"""stub: imported but unused on the EnSRF path (state/ensemble.py:7)"""
class Dataset(object):
    pass
