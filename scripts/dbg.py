import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import chimera_b200 as cb
from oracle import chimera_oracle as orc
from chimera_b200 import synth
from cases import *
g = dict(np.load('tests/golden/golden_models.npz'))
model, kw = MASS_CASES[5]
m = getattr(cb.mass, model)(**kw)
print('norm', m.norm_p_m1, g['mass5_norm'])
p1 = cb.mass.primary_mass_pdf_notnorm(m, g['m1']); r = g['mass5_p1']
bad = np.abs(p1-r) > 1e-11*np.abs(r)
print('p1 bad', g['m1'][bad], p1[bad], r[bad])
p = cb.mass.p_m1m2(m, g['m1'], g['m2']); r = g['mass5_p']
bad = np.abs(p-r) > 1e-10*np.abs(r)
print('p bad', g['m1'][bad], g['m2'][bad], p[bad], r[bad])
cdf = m.cdf_m2_conditioned; r = g['mass5_cdf']
print('cdf maxrel', np.max(np.abs(cdf-r)/np.maximum(np.abs(r),1e-300)))
inj, N_inj = synth.make_injections(30000, seed=9)
sel = cb.selection_function(cb.theta_inj_det(**inj), N_inj, 5.)
pop = cb.population(cb.cosmo.flrw(z_max=5.), cb.mass.plp(), cb.rate.madau_dickinson())
eng = sel._engine(pop)
rows, b = pop.hyper_rows()
lle, part, _ = eng.eval(rows, want_events=False)
print('part', part, 'cfg N_inj', eng.cfg.N_inj, eng.cfg.check_neff, eng.cfg.N_eff, eng.cfg.Tobs)
print(eng.finalize(rows, part, 0))
pop0 = orc.make_pop(orc.make_cosmo("flrw", z_max=5.), orc.make_mass("plp"), orc.make_rate("madau_dickinson"))
print(orc.N_exp(pop0, inj, N_inj, 5.))
w = cb.pop_rate_det(pop, cb.theta_inj_det(**inj))/inj['p_draw']
print('host sums', np.nansum(w), np.sum(w**2))
