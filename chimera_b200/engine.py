"""Thin object wrapper over the C ABI handle (include/chimera_b200.h)."""
import ctypes as C
import numpy as np
from . import _lib


class Engine:
  """Owns one `chb_handle`: device-resident events / pixels / catalogue / injections and the
  batched evaluation entry points."""

  def __init__(self, cfg):
    self.lib = _lib.load()
    self.cfg = cfg
    self.h = C.c_void_p()
    _lib.check(self.lib.chb_create(C.byref(self.h), C.byref(cfg)))
    self.Nev = 0
    self.Nz = 0
    self.P = 0
    self.pixelated = cfg.kind_p_gw != 0
    self._keep = []

  def close(self):
    if getattr(self, "h", None) is not None and self.h.value:
      self.lib.chb_destroy(self.h)
      self.h = C.c_void_p()

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def _chk(self, rc):
    _lib.check(rc, self.h)

  def set_option(self, name, value):
    """Per-handle tuning switch (include/chimera_b200.h: chb_set_option); results do not depend on it."""
    self._chk(self.lib.chb_set_option(self.h, str(name).encode(), float(value)))

  def set_events(self, m1det, m2det, dL, pe_prior, z_grids, ra=None, dec=None):
    m1det, m2det, dL, pe_prior, z_grids, ra, dec = map(_lib.f64, (m1det, m2det, dL, pe_prior, z_grids, ra, dec))
    if dL.ndim != 2 or z_grids.ndim != 2 or z_grids.shape[0] != dL.shape[0]:
      raise ValueError("event arrays must be (Nev, Ns) and z_grids (Nev, Nz)")
    for a in (m1det, m2det, pe_prior, ra, dec):
      if a is not None and a.shape != dL.shape:
        raise ValueError("event arrays must share the shape (Nev, Ns)")
    self.Nev, self.Ns = dL.shape
    self.Nz = z_grids.shape[1]
    self._chk(self.lib.chb_set_events(self.h, self.Nev, self.Ns, self.Nz, _lib.dptr(m1det), _lib.dptr(m2det),
                                      _lib.dptr(dL), _lib.dptr(pe_prior), _lib.dptr(ra), _lib.dptr(dec),
                                      _lib.dptr(z_grids)))

  def set_pixels(self, pixels_opt_nsides, pixels_pe_opt_nside, ra_pix, dec_pix, gw_loc2d_pdf):
    pix, pe = _lib.i64(pixels_opt_nsides), _lib.i64(pixels_pe_opt_nside)
    ra_pix, dec_pix, pdf = map(_lib.f64, (ra_pix, dec_pix, gw_loc2d_pdf))
    if pix.ndim != 2 or pix.shape[0] != self.Nev or ra_pix.shape != pix.shape or dec_pix.shape != pix.shape \
        or pdf.shape != pix.shape:
      raise ValueError("pixel arrays must be (Nev, max_npixels)")
    if pe is not None and pe.shape != (self.Nev, self.Ns):
      raise ValueError("pixels_pe_opt_nside must be (Nev, Ns)")
    self.P = pix.shape[1]
    self._chk(self.lib.chb_set_pixels(self.h, self.P, _lib.iptr(pix), _lib.iptr(pe), _lib.dptr(ra_pix),
                                      _lib.dptr(dec_pix), _lib.dptr(pdf)))

  def set_catalog(self, p_cat, P_compl):
    p_cat = _lib.f64(p_cat)
    P_compl = _lib.f64(np.asarray(P_compl, dtype=np.float64).reshape(self.Nev, self.Nz))
    if p_cat.shape != (self.Nev, self.P, self.Nz):
      raise ValueError("p_cat must be (Nev, max_npixels, Nz)")
    self._chk(self.lib.chb_set_catalog(self.h, _lib.dptr(p_cat), _lib.dptr(P_compl)))

  def set_injections(self, m1det, m2det, dL, p_draw):
    m1det, m2det, dL, p_draw = (np.ascontiguousarray(np.ravel(x), dtype=np.float64) for x in (m1det, m2det, dL, p_draw))
    if not (m1det.shape == m2det.shape == dL.shape == p_draw.shape):
      raise ValueError("injection arrays must share one shape")
    self._chk(self.lib.chb_set_injections(self.h, dL.size, _lib.dptr(m1det), _lib.dptr(m2det), _lib.dptr(dL),
                                          _lib.dptr(p_draw)))

  def eval(self, rows, want_events=True, want_pgw=False):
    """Host entry: rows (n, CHB_NPAR) -> (log_like_evs (n,Nev)|None, partials (n,3), p_gw|None)."""
    rows = _lib.f64(rows)
    n = rows.shape[0]
    lle = np.empty((n, self.Nev)) if (want_events and self.Nev) else None
    part = np.zeros((n, 3))
    pgw = None
    if want_pgw and self.Nev:
      pgw = np.empty((n, self.Nev, self.P, self.Nz) if self.pixelated else (n, self.Nev, self.Nz))
    self._chk(self.lib.chb_eval(self.h, n, _lib.dptr(rows), _lib.dptr(lle), _lib.dptr(part), _lib.dptr(pgw)))
    return lle, part, pgw

  def eval_device(self, d_rows, d_partials, d_log_like=None, stream=None):
    """Device entry for torch tensors (f64, contiguous, on this handle's device); asynchronous."""
    n = d_rows.shape[0]
    self._chk(self.lib.chb_eval_device(self.h, n, d_rows.data_ptr(), d_log_like.data_ptr() if d_log_like is not None else None,
                                       d_partials.data_ptr(), None, stream))

  def numlike_evs(self, n):
    out = np.empty((n, self.Nev))
    self._chk(self.lib.chb_last_numlike_evs(self.h, _lib.dptr(out)))
    return out

  def finalize(self, rows, partials, nev_total):
    rows, partials = _lib.f64(rows), _lib.f64(partials)
    n = rows.shape[0]
    outs = [np.empty(n) for _ in range(5)]
    _lib.check(self.lib.chb_finalize(C.byref(self.cfg), n, int(nev_total), _lib.dptr(rows), _lib.dptr(partials),
                                     *[_lib.dptr(o) for o in outs]))
    return dict(log_like_num=outs[0], log_Nexp=outs[1], log_hyper=outs[2], neff_inj=outs[3], N_exp=outs[4])

  def phase_profile(self, enable=True):
    """Mean SM cycles per CTA in each phase of the numerator kernel (last profiled evaluation)."""
    out = np.zeros(8)
    self._chk(self.lib.chb_phase_profile(self.h, int(enable), _lib.dptr(out)))
    names = ("tables", "zgrid", "reweight", "stats", "kde_integrand", "final")
    return {k: float(v) for k, v in zip(names, out)}

  @property
  def launches(self):
    return int(self.lib.chb_kernel_launch_count(self.h))

  def timings(self):
    t = np.zeros(8)
    self.lib.chb_last_timings(self.h, _lib.dptr(t))
    return dict(tables_ms=t[0], numerator_ms=t[1], selection_ms=t[2], reduce_ms=t[3], zgrid_terms_ms=t[4],
                numerator_kernels_ms=t[5])
