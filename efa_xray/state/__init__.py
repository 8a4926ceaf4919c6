"""alias package, see efa_xray/__init__.py"""
