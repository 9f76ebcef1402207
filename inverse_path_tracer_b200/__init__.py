"""inverse_path_tracer_b200 -- B200-native (sm_100a) drop-in for the rendering hot path of
bblinn2017/inverse_path_tracer: closest-hit queries, the forward path-tracing integrator and the
inverse accumulation / material-adjoint pass, behind the reference's own C ABI.

Layout:
  csrc/      CUDA kernels + host loader + the C ABI (include/ipt_b200.h)
  native.py  ctypes binding of libipt_b200.so (fails loudly when the library is missing)
  ipt_cuda.py  host-side mirror of the reference's ipt_cuda.py entry points
  torch_api.py  torch tensors / autograd.Function over the device-pointer entry points
"""
from .native import IptRenderParams, Scene, default_params, lib, library_path, set_bvh_large_triangle_limit, set_integrator  # noqa: F401
