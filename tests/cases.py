"""Case tables shared by the golden generator's consumers (oracle tests and GPU parity tests).
They restate, as plain data, the model/hyper-point cases `tests/golden/make_golden.py` ran
through the reference."""

COSMO_CASES = [
  ("flrw", dict(H0=70., Om0=0.25)),
  ("flrw", dict(H0=55., Om0=0.4, z_max=5.)),
  ("flrw", dict(H0=67.7, Om0=0.31, w0=-0.9, wa=0.2, Or0=8e-5)),
  ("flrw", dict(H0=80., Om0=0.3, Ok0=0.05)),
  ("flrw", dict(H0=80., Om0=0.3, Ok0=-0.05)),
  ("mg_flrw", dict(H0=70., Om0=0.25, Xi0=1.8, n=1.9)),
  ("mg_flrw", dict(H0=64., Om0=0.3, Xi0=0.6, n=2.5, z_max=5.)),
]
MASS_CASES = [
  ("tpl", {}),
  ("tpl", dict(alpha=1.8, beta=0.3, m_low=4., m_high=60.)),
  ("bpl", {}),
  ("bpl", dict(alpha_1=2.1, alpha_2=4.4, beta=0.5, delta_m=3., break_fraction=0.3, m_low=6., m_high=70.)),
  ("plp", {}),
  ("plp", dict(lambda_peak=0.1, alpha=2.6, beta=-0.4, delta_m=6., mu_g=30., sigma_g=5., m_low=4.2, m_high=95.)),
]
RATE_CASES = [
  ("power_law", {}), ("power_law", dict(gamma=-0.5)),
  ("madau_dickinson", {}), ("madau_dickinson", dict(gamma=1.9, kappa=4.2, zp=1.4)),
  ("trunc_madau_dickinson", dict(zmax=0.9)), ("trunc_power_law", dict(gamma=2.3, zmax=1.1)),
]
LIKE_CASES = {
  # name: (pixel kind, kernel, binning, cosmo model, hyper-points)
  "1d_epan_binned": (None, "epan", True, "flrw",
                     [dict(H0=h) for h in (50., 62., 70., 81., 95.)]),
  "1d_gauss_unbinned": (None, "gauss", False, "flrw",
                        [dict(H0=60., Om0=0.2), dict(H0=70., Om0=0.25), dict(H0=78., Om0=0.4)]),
  "1d_gauss_binned_mg": (None, "gauss", True, "mg_flrw",
                         [dict(H0=70., Xi0=1.0, n=0.), dict(H0=66., Xi0=1.6, n=1.9, alpha=3.0, mu_g=32., gamma=2.2),
                          dict(H0=74., Xi0=0.7, n=2.4, beta=0.8, delta_m=5.5, m_low=4.8, m_high=90., sigma_g=4.2,
                               lambda_peak=0.06, kappa=3.5, zp=1.8)]),
  "1d_epan_unbinned": (None, "epan", False, "flrw", [dict(H0=65.), dict(H0=75., Om0=0.3)]),
  "approx_gauss_unbinned": ("approximate", "gauss", False, "flrw",
                            [dict(H0=58., Om0=0.22), dict(H0=70., Om0=0.25), dict(H0=84., Om0=0.33)]),
  "approx_epan_binned": ("approximate", "epan", True, "flrw", [dict(H0=64.), dict(H0=70.), dict(H0=77.)]),
  "marg_binned": ("marginalized", "epan", True, "flrw", [dict(H0=61.), dict(H0=70.), dict(H0=88.)]),
  "marg_unbinned": ("marginalized", "epan", False, "flrw", [dict(H0=66.), dict(H0=73., Om0=0.28)]),
  "full_gauss": ("full", "gauss", False, "flrw", [dict(H0=63.), dict(H0=70.), dict(H0=79., Om0=0.3)]),
}
GOLDEN_NUM_BINS = 40

# Round 2: the model / option matrix the fp32 mode had never been compared on:
# name: dict(kind, kernel, binning, bw, cosmo=(model, kwargs), mass=(model, kwargs), rate=(model, kwargs), hypers)
LIKE_CASES2 = {
  "1d_gauss_tpl_pl": dict(kind=None, kernel="gauss", binning=False, bw=None, cosmo=("flrw", {}), mass=("tpl", {}),
                          rate=("power_law", {}),
                          hypers=[dict(H0=62., alpha=3.0), dict(H0=75., beta=1.6, gamma=1.2), dict(H0=70., alpha=12.0)]),
  "1d_gauss_bpl_tpl_rate": dict(kind=None, kernel="gauss", binning=False, bw=None, cosmo=("flrw", {}), mass=("bpl", {}),
                                rate=("trunc_power_law", dict(zmax=2.5)),
                                hypers=[dict(H0=66.), dict(H0=74., alpha_1=2.0, alpha_2=4.4, break_fraction=0.3, delta_m=3.5),
                                        dict(H0=58., Om0=0.35, beta=0.4, m_low=4.2, m_high=95., gamma=2.4, zmax=0.9)]),
  "1d_gauss_silverman": dict(kind=None, kernel="gauss", binning=False, bw="silverman", cosmo=("flrw", {}), mass=("plp", {}),
                             rate=("madau_dickinson", {}), hypers=[dict(H0=64.), dict(H0=77., Om0=0.31)]),
  "1d_gauss_scalar_bw": dict(kind=None, kernel="gauss", binning=False, bw=0.3, cosmo=("flrw", {}), mass=("plp", {}),
                             rate=("madau_dickinson", {}), hypers=[dict(H0=64.), dict(H0=77., Om0=0.31)]),
  "1d_gauss_curved_w0wa": dict(kind=None, kernel="gauss", binning=False, bw=None, cosmo=("flrw", {}), mass=("plp", {}),
                               rate=("madau_dickinson", {}),
                               hypers=[dict(H0=70., Ok0=0.05), dict(H0=70., Ok0=-0.05), dict(H0=67., Om0=0.31, w0=-0.9, wa=0.2),
                                       dict(H0=72., Om0=0.28, Ok0=0.03, w0=-1.1, wa=-0.3, Or0=8e-5)]),
  "1d_epan_binned_bpl_silverman": dict(kind=None, kernel="epan", binning=True, bw="silverman", cosmo=("flrw", {}),
                                       mass=("bpl", {}), rate=("trunc_madau_dickinson", dict(zmax=1.5)),
                                       hypers=[dict(H0=63.), dict(H0=76., alpha_1=1.9, kappa=3.5)]),
  "1d_epan_unbinned_scalar_tpl": dict(kind=None, kernel="epan", binning=False, bw=0.4, cosmo=("flrw", {}), mass=("tpl", {}),
                                      rate=("power_law", {}), hypers=[dict(H0=68.), dict(H0=79., alpha=2.2)]),
  "approx_gauss_tpl_mg": dict(kind="approximate", kernel="gauss", binning=False, bw=None, cosmo=("mg_flrw", {}),
                              mass=("tpl", {}), rate=("trunc_madau_dickinson", dict(zmax=1.6)),
                              hypers=[dict(H0=70., Xi0=1.0, n=0.), dict(H0=66., Xi0=1.7, n=1.9, alpha=2.8),
                                      dict(H0=75., Xi0=0.7, n=2.5, beta=0.6)]),
  "marg_binned_bpl": dict(kind="marginalized", kernel="epan", binning=True, bw=None, cosmo=("flrw", {}), mass=("bpl", {}),
                          rate=("power_law", {}), hypers=[dict(H0=63.), dict(H0=78., alpha_2=4.0, gamma=2.5)]),
  "marg_unbinned_silverman_tpl": dict(kind="marginalized", kernel="epan", binning=False, bw="silverman", cosmo=("flrw", {}),
                                      mass=("tpl", {}), rate=("madau_dickinson", {}), hypers=[dict(H0=65.), dict(H0=74., Ok0=0.04)]),
  "full_gauss_tpl_pl": dict(kind="full", kernel="gauss", binning=False, bw=None, cosmo=("flrw", {}), mass=("tpl", {}),
                            rate=("trunc_power_law", dict(zmax=1.4)), hypers=[dict(H0=64.), dict(H0=77., alpha=2.9, gamma=1.4)]),
  "full_gauss_silverman_bpl": dict(kind="full", kernel="gauss", binning=False, bw="silverman", cosmo=("flrw", {}),
                                   mass=("bpl", {}), rate=("madau_dickinson", {}), hypers=[dict(H0=70.), dict(H0=81., w0=-0.9)]),
}
SEL_BPL_MG_HYPERS = [dict(H0=60., Xi0=0.7, n=1.9), dict(H0=70., Xi0=1.0, n=0.),
                     dict(H0=82., Xi0=1.8, n=2.3, alpha_1=2.0, break_fraction=0.3)]
