"""ctypes binding of ``libnl_b200.so`` (C ABI declared in ``include/nl_b200.h``).

The library is built in-tree by ``neurallaplace_b200/csrc/build.sh`` (or
``__graft_entry__.build()``).  There is NO fallback: if the shared object is missing or a CUDA
device is absent, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# NL_B200_LIB: another build of the same library (the -DNL_TC_PROFILE variant of scripts/gpu_phase_profile.sh)
LIB_PATH = os.environ.get("NL_B200_LIB") or os.path.join(_HERE, "lib", "libnl_b200.so")

NL_F32, NL_F64 = 0, 1
NL_FOURIER, NL_AW, NL_DEHOOG = 0, 1, 2
NL_T_SHARED, NL_T_PER_ROW = 0, 1
NL_HEAD_SPHERE, NL_HEAD_RAW = 0, 1
NL_F_SPHERE_PLANAR, NL_F_COMPLEX_PLANAR, NL_F_COMPLEX_INTERLEAVED = 0, 1, 2
NL_ERR_UNSUPPORTED = -3

# every symbol include/nl_b200.h declares (tests check the .so exports all of them)
EXPORTS = (
    "nl_abi_version",
    "nl_strerror",
    "nl_last_cuda_error",
    "nl_workspace_bytes",
    "nl_selected_impl",
    "nl_last_launch_count",
    "nl_reconstruct_forward",
    "nl_reconstruct_backward",
    "nl_saved_bytes",
    "nl_reconstruct_forward_save",
    "nl_reconstruct_backward_saved",
    "nl_query_points",
    "nl_line_integrate_forward",
    "nl_line_integrate_backward",
    "nl_complex_to_sphere",
    "nl_sphere_to_complex",
    "nl_line_integrate_scratch_bytes",
    "nl_set_scratch",
    "nl_dehoog_fixed_scratch_bytes",
    "nl_dehoog_fixed_forward",
    "nl_dehoog_fixed_backward",
    "nl_clip_adam_scratch_bytes",
    "nl_clip_adam_step",
)


class NlDesc(ctypes.Structure):
  _fields_ = [
      ("B", ctypes.c_int32),
      ("T", ctypes.c_int32),
      ("K", ctypes.c_int32),
      ("H", ctypes.c_int32),
      ("D", ctypes.c_int32),
      ("n", ctypes.c_int32),
      ("family", ctypes.c_int32),
      ("t_mode", ctypes.c_int32),
      ("head", ctypes.c_int32),
      ("dtype", ctypes.c_int32),
      ("impl", ctypes.c_int32),
      ("reserved", ctypes.c_int32),
      ("alpha", ctypes.c_double),
      ("log_tol", ctypes.c_double),
      ("scale", ctypes.c_double),
  ]


class NlError(RuntimeError):
  pass


_lib = None
_lock = threading.Lock()


def load():
  """Load the shared library (once).  Raises if it has not been built."""
  global _lib
  if _lib is not None:
    return _lib
  with _lock:
    if _lib is not None:
      return _lib
    if not os.path.exists(LIB_PATH):
      raise NlError(f"{LIB_PATH} not found: build it with neurallaplace_b200/csrc/build.sh "
                    "(or __graft_entry__.build()); there is no CPU / PyTorch fallback")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
    dp = ctypes.POINTER(NlDesc)
    pp = ctypes.POINTER(ctypes.c_void_p)
    lib.nl_abi_version.restype = i32
    lib.nl_strerror.restype = ctypes.c_char_p
    lib.nl_strerror.argtypes = [i32]
    lib.nl_last_cuda_error.restype = ctypes.c_char_p
    lib.nl_last_launch_count.restype = i32
    lib.nl_workspace_bytes.restype = sz
    lib.nl_workspace_bytes.argtypes = [dp, i32]
    lib.nl_selected_impl.restype = i32
    lib.nl_selected_impl.argtypes = [dp, i32]
    lib.nl_reconstruct_forward.restype = i32
    lib.nl_reconstruct_forward.argtypes = [dp, vp, vp, pp, vp, vp, vp, vp, sz, vp]
    lib.nl_reconstruct_backward.restype = i32
    lib.nl_reconstruct_backward.argtypes = [dp, vp, vp, pp, vp, vp, vp, pp, vp, sz, vp]
    lib.nl_saved_bytes.restype = sz
    lib.nl_saved_bytes.argtypes = [dp]
    lib.nl_reconstruct_forward_save.restype = i32
    lib.nl_reconstruct_forward_save.argtypes = [dp, vp, vp, pp, vp, vp, vp, vp, sz, vp, sz, vp]
    lib.nl_reconstruct_backward_saved.restype = i32
    lib.nl_reconstruct_backward_saved.argtypes = [dp, vp, vp, pp, vp, vp, vp, vp, sz, pp, vp, sz, vp]
    lib.nl_query_points.restype = i32
    lib.nl_query_points.argtypes = [dp, i64, vp, vp, vp, vp, vp, vp]
    lib.nl_line_integrate_forward.restype = i32
    lib.nl_line_integrate_forward.argtypes = [dp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.nl_line_integrate_backward.restype = i32
    lib.nl_line_integrate_backward.argtypes = [dp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.nl_complex_to_sphere.restype = i32
    lib.nl_complex_to_sphere.argtypes = [i32, i64, vp, vp, vp, vp, vp]
    lib.nl_sphere_to_complex.restype = i32
    lib.nl_sphere_to_complex.argtypes = [i32, i64, vp, vp, vp, vp, vp]
    lib.nl_line_integrate_scratch_bytes.restype = sz
    lib.nl_line_integrate_scratch_bytes.argtypes = [dp, i32]
    lib.nl_set_scratch.restype = None
    lib.nl_set_scratch.argtypes = [vp, sz]
    f64 = ctypes.c_double
    lib.nl_dehoog_fixed_scratch_bytes.restype = sz
    lib.nl_dehoog_fixed_scratch_bytes.argtypes = [i32, i64, i32, i64, i32]
    lib.nl_dehoog_fixed_forward.restype = i32
    lib.nl_dehoog_fixed_forward.argtypes = [i32, i64, i32, i64, vp, vp, f64, f64, vp, vp, sz, vp]
    lib.nl_dehoog_fixed_backward.restype = i32
    lib.nl_dehoog_fixed_backward.argtypes = [i32, i64, i32, i64, vp, vp, f64, f64, vp, vp, vp, sz, vp]
    lib.nl_clip_adam_scratch_bytes.restype = sz
    lib.nl_clip_adam_step.restype = i32
    lib.nl_clip_adam_step.argtypes = [i32, i64, vp, vp, vp, vp, f64, f64, f64, f64, i64, f64, i32, vp, sz, vp]
    _lib = lib
  return _lib


def check(status):
  if status != 0:
    lib = load()
    msg = lib.nl_strerror(status).decode()
    if status == -5:
      msg += " (" + lib.nl_last_cuda_error().decode() + ")"
    raise NlError(f"libnl_b200: {msg} [status {status}]")


def require_cuda():
  if not torch.cuda.is_available():
    raise NlError("neurallaplace_b200 needs a CUDA device: the hot path is hand-written sm_100a CUDA and "
                  "there is no CPU fallback")
  load()


def dtype_code(dtype):
  if dtype == torch.float32:
    return NL_F32
  if dtype == torch.float64:
    return NL_F64
  raise TypeError(f"unsupported dtype {dtype}: the reconstruction path computes in float32 or float64")


def ptr(t):
  return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)  # pylint: disable=protected-access


def stream_ptr(device=None):
  """Current torch stream of ``device`` (default: the current device) as the ABI's ``void* cuda_stream``."""
  if _raw_stream is not None:  # one C call instead of building a torch.cuda.Stream object (hot path of small steps)
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    return ctypes.c_void_p(_raw_stream(idx))
  return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _Same:
  """No-op context: the tensors' device is already the current one (the common case; saves two driver calls)."""

  def __enter__(self):
    return self

  def __exit__(self, *a):
    return False


_SAME = _Same()


def on(device):
  """Context that makes ``device`` the current CUDA device for a C-ABI call: the library launches on the
  current device (kernel attributes, SM count, launches), so every call site enters this with the device of
  its tensors and passes ``stream_ptr(device)``."""
  if device.index is None or device.index == torch.cuda.current_device():
    return _SAME
  return torch.cuda.device(device)


def ptr_array(tensors):
  arr = (ctypes.c_void_p * len(tensors))()
  for i, t in enumerate(tensors):
    arr[i] = t.data_ptr()
  return arr
