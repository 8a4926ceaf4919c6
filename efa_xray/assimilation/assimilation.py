from efa_xray_b200.assimilation.assimilation import (Assimilation, update, ObTimeOutsideState,  # noqa: F401
                                                       randomize_obs_order)
