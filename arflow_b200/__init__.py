"""arflow_b200 — B200-native (sm_100a) kernels for the unsupervised-flow hot path of deu439/ARFlow.

Module names mirror the reference so that `from utils.warp_utils import flow_warp` becomes
`from arflow_b200.warp_utils import flow_warp` and nothing else changes:

    arflow_b200.correlation      models/correlation_package/correlation.py, models/correlation_native.py
    arflow_b200.warp_utils       utils/warp_utils.py
    arflow_b200.uflow_utils      utils/uflow_utils.py
"""
__version__ = "0.1.0"
