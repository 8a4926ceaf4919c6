from efa_xray_b200.observation.observation import Observation, gaspari_cohn, haversine  # noqa: F401
