"""pod_uqnn_b200 -- B200-native (sm_100a) hot path of POD-UQNN behind the reference's API.

    from pod_uqnn_b200.pod import perform_pod, perform_fast_pod
    from pod_uqnn_b200.acceleration import loop_u, loop_u_t
    from pod_uqnn_b200.podnnmodel import PodnnModel
    from pod_uqnn_b200.varneuralnetwork import VarNeuralNetwork

Everything computes in libpodb200.so (CUDA, built in-tree); there is no CPU fallback.
"""
from ._lib import Context, DeviceArray, default_context, lib, LIB_PATH  # noqa: F401

__all__ = ["Context", "DeviceArray", "default_context", "lib", "LIB_PATH"]
