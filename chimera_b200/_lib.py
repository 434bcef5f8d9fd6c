"""ctypes binding of libchimera_b200.so (C ABI: include/chimera_b200.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and
without a CUDA device every compute entry point returns CHB_ERR_CUDA, which `check` turns
into a RuntimeError."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHB_LIB") or os.path.join(_HERE, "libchimera_b200.so")     # CHB_LIB: A/B builds

CHB_ABI_VERSION = 2
CHB_NPAR = 32
OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = 0, -1, -2, -3, -4

COSMO_IDS = {"flrw": 0, "mg_flrw": 1}
MASS_IDS = {"truncated_power_law": 0, "broken_power_law": 1, "power_law_plus_peak": 2}
RATE_IDS = {"power_law": 0, "madau_dickinson": 1, "trunc_madau_dickinson": 2, "trunc_power_law": 3}
KERNEL_IDS = {"epan": 0, "gauss": 1}
KIND_IDS = {None: 0, "approximate": 1, "marginalized": 2, "full": 3}
FP_IDS = {"fp64": 0, "fp32": 1}

# hyper-row slots (enum in chimera_b200.h)
SLOT = dict(H0=0, Om0=1, Ok0=2, Or0=3, w0=4, wa=5, Xi0=6, n=7, z_max=8,
            m_low=9, m_high=10, alpha=11, alpha_1=11, beta=12, delta_m=13, alpha_2=14, break_fraction=15,
            lambda_peak=16, mu_g=17, sigma_g=18, gamma=21, kappa=22, zp=23, zmax=24, R0=25)

F_E_AT_Z, F_DL_AT_Z, F_Z_FROM_DGW, F_DDLDZ_AT_Z, F_DVCDZ_AT_Z, F_VC_AT_Z, F_DCT_AT_Z, F_P_M1M2, \
    F_P_M1_NOTNORM, F_MERGER_RATE, F_POP_RATE_DET_INJ = range(11)


class chb_config(C.Structure):
  _fields_ = [
    ("abi_version", C.c_int32), ("device", C.c_int32), ("fp_mode", C.c_int32),
    ("cosmo_model", C.c_int32), ("mass_model", C.c_int32), ("rate_model", C.c_int32),
    ("cosmo_grid_res", C.c_int32), ("mass_grid_res", C.c_int32),
    ("kind_p_gw", C.c_int32), ("kernel", C.c_int32), ("bw_method", C.c_int32), ("bw_value", C.c_double),
    ("use_cut_grid", C.c_int32), ("cut_grid", C.c_double), ("binning", C.c_int32), ("num_bins", C.c_int32),
    ("pe_neff", C.c_double), ("scale_free", C.c_int32), ("Tobs", C.c_double),
    ("catalog_kind", C.c_int32), ("compl_z_lo", C.c_double), ("compl_z_hi", C.c_double),
    ("N_inj", C.c_double), ("check_neff", C.c_int32), ("N_eff", C.c_double),
  ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_hp = C.c_void_p

EXPORTS = {
  "chb_abi_version": (C.c_int, []),
  "chb_device_count": (C.c_int, []),
  "chb_last_error": (C.c_char_p, [_hp]),
  "chb_create": (C.c_int, [C.POINTER(_hp), C.POINTER(chb_config)]),
  "chb_destroy": (None, [_hp]),
  "chb_set_option": (C.c_int, [_hp, C.c_char_p, C.c_double]),
  "chb_set_events": (C.c_int, [_hp, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
  "chb_set_pixels": (C.c_int, [_hp, C.c_int64, _ip, _ip, _dp, _dp, _dp]),
  "chb_set_catalog": (C.c_int, [_hp, _dp, _dp]),
  "chb_set_injections": (C.c_int, [_hp, C.c_int64, _dp, _dp, _dp, _dp]),
  "chb_eval": (C.c_int, [_hp, C.c_int64, _dp, _dp, _dp, _dp]),
  "chb_eval_device": (C.c_int, [_hp, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
  "chb_last_numlike_evs": (C.c_int, [_hp, _dp]),
  "chb_finalize": (C.c_int, [C.POINTER(chb_config), C.c_int64, C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
  "chb_model_eval": (C.c_int, [C.POINTER(chb_config), C.c_int, _dp, C.c_int64, _dp, _dp, _dp, _dp]),
  "chb_model_tables": (C.c_int, [C.POINTER(chb_config), _dp, _dp, _dp, _dp, _dp, _dp]),
  "chb_kernel_launch_count": (C.c_int64, [_hp]),
  "chb_last_timings": (C.c_int, [_hp, _dp]),
  "chb_phase_profile": (C.c_int, [_hp, C.c_int, _dp]),
  "chb_mufu_peak": (C.c_int, [C.c_int, C.c_double, _dp]),
  "chb_healpix_ang2pix_ring": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _dp, _dp, _ip]),
  "chb_healpix_pix2ang_ring": (C.c_int, [C.c_int, C.c_int64, C.c_int64, _ip, _dp, _dp]),
  "chb_pixelize_samples": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, _ip, _dp, _dp, _ip, _dp, _dp, _ip, _dp]),
  "chb_precompute_p_cat": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _ip, _ip, C.POINTER(C.c_int32),
                                     C.c_int64, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
}

_lib = None


def load():
  """Load the shared library (once) and declare every prototype of include/chimera_b200.h."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} is missing: build it with `python -m chimera_b200.build` "
                      "(nvcc, sm_100a). chimera_b200 has no CPU fallback.")
  lib = C.CDLL(LIB_PATH)
  for name, (res, args) in EXPORTS.items():
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = args
  if lib.chb_abi_version() != CHB_ABI_VERSION:
    raise ImportError("libchimera_b200.so ABI version mismatch: rebuild the library")
  _lib = lib
  return lib


def dptr(a):
  return None if a is None else a.ctypes.data_as(_dp)


def iptr(a):
  return None if a is None else a.ctypes.data_as(_ip)


def f64(a):
  return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def i64(a):
  return None if a is None else np.ascontiguousarray(a, dtype=np.int64)


def check(rc, handle=None):
  if rc == OK:
    return
  msg = load().chb_last_error(handle)
  msg = msg.decode() if msg else "unknown error"
  if rc in (ERR_INVALID, ERR_UNSUPPORTED):
    raise ValueError(msg)
  raise RuntimeError(f"chimera_b200 ({rc}): {msg}")


def device_count():
  return int(load().chb_device_count())
