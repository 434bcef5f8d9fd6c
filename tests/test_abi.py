"""CPU checks of the C-ABI boundary: the library loads, exports every symbol the header declares,
fails loudly without a GPU, validates configurations, and its host epilogue `chb_finalize`
reproduces selection_function.N_exp / compute_log_hyperlike arithmetic.  No GPU compute here."""
import ctypes as C
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
  import __graft_entry__ as ge
  ge.build()
  from chimera_b200 import _lib
  return _lib.load()


def test_header_symbols_exported(lib):
  from chimera_b200 import _lib
  hdr = open(os.path.join(ROOT, "include", "chimera_b200.h")).read()
  declared = set(re.findall(r"\b(chb_[a-z0-9_]+)\s*\(", hdr))
  assert declared, "no declarations found in the header"
  for name in sorted(declared):
    assert hasattr(lib, name), f"{name} declared in chimera_b200.h but not exported"
  assert declared == set(_lib.EXPORTS), "ctypes prototype table and header disagree"
  assert lib.chb_abi_version() == _lib.CHB_ABI_VERSION


def test_config_struct_layout():
  """ctypes mirror of chb_config: compile a probe with gcc and compare sizeof/offsets."""
  import subprocess, tempfile
  from chimera_b200 import _lib
  src = r'''
#include <stdio.h>
#include <stddef.h>
#include "chimera_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(chb_config), offsetof(chb_config, bw_value),
  offsetof(chb_config, cut_grid), offsetof(chb_config, pe_neff), offsetof(chb_config, N_inj), offsetof(chb_config, N_eff)); return 0; }
'''
  with tempfile.TemporaryDirectory() as d:
    open(os.path.join(d, "p.c"), "w").write(src)
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
    out = subprocess.check_output([os.path.join(d, "p")]).split()
  got = [int(x) for x in out]
  cfg = _lib.chb_config
  want = [C.sizeof(cfg), cfg.bw_value.offset, cfg.cut_grid.offset, cfg.pe_neff.offset, cfg.N_inj.offset, cfg.N_eff.offset]
  assert got == want


def _cfg(**kw):
  from chimera_b200.population._base import model_config
  from chimera_b200 import cosmo, mass, rate
  return model_config(cosmo.flrw(), mass.plp(), rate.madau_dickinson(), **kw)


def test_no_gpu_fails_loudly(lib):
  from chimera_b200 import _lib
  if lib.chb_device_count() > 0:
    pytest.skip("a GPU is present")
  h = C.c_void_p()
  cfg = _cfg()
  rc = lib.chb_create(C.byref(h), C.byref(cfg))
  assert rc == _lib.ERR_CUDA and not h.value
  assert b"no CPU fallback" in lib.chb_last_error(None)
  with pytest.raises(RuntimeError):
    import chimera_b200 as cb
    cb.cosmo.dL_at_z(cb.cosmo.flrw(), np.array([0.1]))
  out = np.zeros(1)
  assert lib.chb_mufu_peak(0, 0.01, _lib.dptr(out)) == _lib.ERR_CUDA
  # the setup-side entry points (HEALPix, pixelisation, p_cat) have no host fallback either
  with pytest.raises(RuntimeError):
    cb.sky.ang2pix(8, np.array([0.3]), np.array([1.0]))
  with pytest.raises(RuntimeError):
    cb.sky.pix2ang(8, np.array([5]))
  th = cb.theta_pe_det(dL=np.ones((1, 4)), ra=np.full((1, 4), 0.1), dec=np.full((1, 4), 0.2))
  with pytest.raises(RuntimeError):
    cb.pixelize_gw_catalog(th, [8], 4, 0.9)


def test_setup_entry_points_validate_before_touching_the_device(lib):
  """nside must be a power of two and angles in range (healpy raises ValueError): checked on the host first."""
  import chimera_b200 as cb
  with pytest.raises(ValueError):
    cb.sky.ang2pix(12, np.array([0.3]), np.array([1.0]))
  with pytest.raises(ValueError):
    cb.sky.ang2pix(8, np.array([-0.3]), np.array([1.0]))
  with pytest.raises(ValueError):
    cb.sky.pix2ang(8, np.array([12 * 64]))
  with pytest.raises(NotImplementedError):
    cb.sky.ang2pix(8, 0.1, 0.1, nest=True)


def test_config_validation(lib):
  from chimera_b200 import _lib
  h = C.c_void_p()
  for bad in (dict(kind_p_gw=7), dict(kernel=5), dict(bw_method=9), dict(abi_version=99), dict(fp_mode=3),
              dict(binning=1, num_bins=0), dict(cosmo_grid_res=2)):
    cfg = _cfg(**bad)
    assert lib.chb_create(C.byref(h), C.byref(cfg)) == _lib.ERR_INVALID, bad
  cfg = _cfg(kind_p_gw=3, use_cut_grid=0)
  assert lib.chb_create(C.byref(h), C.byref(cfg)) == _lib.ERR_UNSUPPORTED
  assert b"cut_grid" in lib.chb_last_error(None)


def _finalize(lib, cfg, rows, part, nev):
  from chimera_b200 import _lib
  n = rows.shape[0]
  outs = [np.empty(n) for _ in range(5)]
  rc = lib.chb_finalize(C.byref(cfg), n, nev, _lib.dptr(rows), _lib.dptr(part), *[_lib.dptr(o) for o in outs])
  assert rc == 0
  return outs


def test_finalize_matches_reference_formulas(lib):
  """selection_function.py:37-47 and likelihood.py:299-316 restated with NumPy."""
  from chimera_b200.population._base import base_rows
  rng = np.random.default_rng(0)
  n, nev, Ninj = 6, 37, 5.0e5
  rows = base_rows(n, R0=rng.uniform(5, 30, n))
  part = np.stack([rng.uniform(-300, -100, n), rng.uniform(2e6, 5e6, n), rng.uniform(1e9, 5e9, n)], axis=1)
  part[1, 0] = -np.finfo(np.float64).max
  part[2, 0] = -np.inf
  part[3, 2] = 1e15            # variance so large that neff < N_eff -> N_exp = 0
  for scale_free in (1, 0):
    cfg = _cfg(scale_free=scale_free, Tobs=2.5, N_inj=Ninj, check_neff=1, N_eff=5.0)
    lnum, lnexp, lh, neff, nexp = _finalize(lib, cfg, rows, part, nev)
    with np.errstate(all="ignore"):
      xi = part[:, 1] / Ninj
      var = part[:, 2] / Ninj ** 2 - xi ** 2 / Ninj
      rneff = xi ** 2 / var
      rnexp = np.where(rneff < 5.0, 0.0, 2.5 * xi)
      if scale_free:
        rnum = part[:, 0]
        rlh = rnum - nev * np.log(rnexp)
      else:
        rnum = part[:, 0] + nev * np.log(rows[:, 25] * 2.5)
        rlh = rnum - rnexp
    np.testing.assert_allclose(neff, rneff, rtol=1e-14)
    np.testing.assert_array_equal(nexp, rnexp)
    np.testing.assert_allclose(lnum, rnum, rtol=1e-15)
    np.testing.assert_allclose(lnexp, np.log(rnexp), rtol=1e-15)
    np.testing.assert_allclose(lh, rlh, rtol=1e-15)
    assert rnexp[3] == 0.0 and (lh[3] == np.inf if scale_free else True)   # the reference's +inf quirk
  cfg = _cfg(scale_free=1, Tobs=1.0, N_inj=Ninj, check_neff=0)
  _, _, _, neff, nexp = _finalize(lib, cfg, rows, part, nev)
  assert np.all(np.isnan(neff)) and nexp[3] > 0                              # N_eff=None: no gate
