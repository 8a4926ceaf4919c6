from efa_xray_b200.assimilation.ensrf import EnSRF  # noqa: F401
