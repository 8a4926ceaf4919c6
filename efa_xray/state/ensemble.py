from efa_xray_b200.state.ensemble import *  # noqa: F401,F403
from efa_xray_b200.state.ensemble import EnsembleState  # noqa: F401
