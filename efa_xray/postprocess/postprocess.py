"""alias module, see efa_xray/__init__.py"""
from efa_xray_b200.postprocess.postprocess import obs_assimilation_statistics, ob_estimates  # noqa: F401
