"""world_size-2 gloo test of the multi-process host logic: contiguous event /
injection shards, one SUM all-reduce of the (n_hyper, 3) partials, all-gather of per-event values,
then the host epilogue `chb_finalize` -- must reproduce the single-process result.  The per-rank
partials come from the oracle here (no GPU in this container); on the GPU box the same
`parallel.*` functions carry the CUDA kernels' partials over NCCL."""
import os
import socket
import numpy as np
import pytest


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _partials_oracle(ev, zg, inj, N_inj, hypers, lo, hi, ilo, ihi):
  from oracle import chimera_oracle as orc
  pop0 = orc.make_pop(orc.make_cosmo("flrw", z_max=5.), orc.make_mass("plp"), orc.make_rate("madau_dickinson"))
  opts = orc.make_opts(None, "epan", None, 2.0, True, 40, 2.0)
  evs = {k: v[lo:hi] for k, v in ev.items()}
  injs = {k: v[ilo:ihi] for k, v in inj.items()}
  lle, part = [], []
  with np.errstate(all="ignore"):
    for hl in hypers:
      pop = orc.pop_update(pop0, **hl)
      l = np.nan_to_num(np.log(orc.numlike_evs(pop, evs, zg[lo:hi], opts)), nan=-np.inf) if hi > lo else np.zeros(0)
      w = orc.pop_rate_det_inj(pop, injs) / injs["p_draw"]
      lle.append(l)
      part.append([np.sum(l), np.nansum(w), np.sum(w ** 2)])
  return np.array(lle), np.array(part)


def _worker(rank, world, port, q):
  import ctypes as C
  import torch.distributed as dist
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from chimera_b200 import parallel, _lib, synth
    from chimera_b200.population._base import base_rows, model_config
    import chimera_b200 as cb
    assert parallel.dist_info() == (rank, world)
    ev = synth.make_events(7, 300, seed=5)
    ev = {k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior")}
    zg = synth.make_z_grids(ev["dL"], 60, H0_prior=(40., 120.))
    inj, N_inj = synth.make_injections(2001, seed=6)
    hypers = [dict(H0=60.), dict(H0=70.), dict(H0=83.)]
    lo, hi = parallel.shard_bounds(7, rank, world)
    ilo, ihi = parallel.shard_bounds(2001, rank, world)
    lle, part = _partials_oracle(ev, zg, inj, N_inj, hypers, lo, hi, ilo, ihi)
    counts = parallel.allgather_counts(hi - lo)
    part = parallel.allreduce_partials(part)
    lle_all = parallel.allgather_events(lle, counts)
    pop = cb.population(cb.cosmo.flrw(z_max=5.), cb.mass.plp(), cb.rate.madau_dickinson()).update(H0=[60., 70., 83.])
    rows, _ = pop.hyper_rows()
    cfg = model_config(pop.cosmo, pop.mass, pop.rate, N_inj=float(N_inj), check_neff=1, N_eff=5.0)
    outs = [np.empty(3) for _ in range(5)]
    rc = _lib.load().chb_finalize(C.byref(cfg), 3, 7, _lib.dptr(rows), _lib.dptr(np.ascontiguousarray(part)),
                                  *[_lib.dptr(o) for o in outs])
    assert rc == 0
    q.put((rank, counts, lle_all, outs[2]))
  finally:
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
  import torch.multiprocessing as mp
  import __graft_entry__ as ge
  ge.build()
  from chimera_b200 import synth
  from oracle import chimera_oracle as orc
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=240) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  # single-process reference
  ev = synth.make_events(7, 300, seed=5)
  ev = {k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior")}
  zg = synth.make_z_grids(ev["dL"], 60, H0_prior=(40., 120.))
  inj, N_inj = synth.make_injections(2001, seed=6)
  pop0 = orc.make_pop(orc.make_cosmo("flrw", z_max=5.), orc.make_mass("plp"), orc.make_rate("madau_dickinson"))
  opts = orc.make_opts(None, "epan", None, 2.0, True, 40, 2.0)
  ref = [orc.compute_all(pop0, ev, zg, opts, inj, N_inj, 5., None, H0=h) for h in (60., 70., 83.)]
  for rank, counts, lle_all, lh in res:
    assert counts == [4, 3]
    for i in range(3):
      np.testing.assert_allclose(lle_all[i], ref[i][0], rtol=1e-13)
      np.testing.assert_allclose(lh[i], ref[i][3], rtol=1e-12)


def _worker_2d(rank, world, port, q):
  """4 ranks = 2 event shards x 2 hyper groups (the reference's 'both' scheme, CHIMERA/parallel.py:132-229)."""
  import torch.distributed as dist
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from chimera_b200 import parallel, synth
    ev = synth.make_events(7, 300, seed=5)
    ev = {k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior")}
    zg = synth.make_z_grids(ev["dL"], 60, H0_prior=(40., 120.))
    inj, N_inj = synth.make_injections(2001, seed=6)
    hypers = [dict(H0=60.), dict(H0=70.), dict(H0=83.)]
    ei, E, gi = parallel.grid_coords(rank, world, hyper_groups=2)
    lo, hi = parallel.shard_bounds(7, ei, E)
    ilo, ihi = parallel.shard_bounds(2001, ei, E)
    hlo, hhi = parallel.shard_bounds(len(hypers), gi, 2)
    part = np.zeros((len(hypers), 3))
    lle = np.zeros((len(hypers), hi - lo))
    l, p = _partials_oracle(ev, zg, inj, N_inj, hypers[hlo:hhi], lo, hi, ilo, ihi)
    part[hlo:hhi] = p
    lle[hlo:hhi] = l
    part = parallel.allreduce_partials(part)
    counts = [b - a for a, b in (parallel.shard_bounds(7, r, E) for r in range(E))]
    lle_all = parallel.allgather_events(lle, counts * 2)
    lle_all = lle_all[:, :7] + lle_all[:, 7:]
    q.put((rank, (ei, E, gi), part, lle_all))
  finally:
    dist.destroy_process_group()


def test_four_rank_2d_sharding_matches_single_process():
  import torch.multiprocessing as mp
  from chimera_b200 import synth, parallel
  assert [parallel.grid_coords(r, 8, 2) for r in range(8)] == [(0, 4, 0), (1, 4, 0), (2, 4, 0), (3, 4, 0),
                                                                (0, 4, 1), (1, 4, 1), (2, 4, 1), (3, 4, 1)]
  with pytest.raises(ValueError):
    parallel.grid_coords(0, 8, 3)
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker_2d, args=(r, 4, port, q)) for r in range(4)]
  for p in procs:
    p.start()
  res = [q.get(timeout=240) for _ in procs]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  ev = synth.make_events(7, 300, seed=5)
  ev = {k: ev[k] for k in ("m1det", "m2det", "dL", "pe_prior")}
  zg = synth.make_z_grids(ev["dL"], 60, H0_prior=(40., 120.))
  inj, N_inj = synth.make_injections(2001, seed=6)
  hypers = [dict(H0=60.), dict(H0=70.), dict(H0=83.)]
  lle_ref, part_ref = _partials_oracle(ev, zg, inj, N_inj, hypers, 0, 7, 0, 2001)
  for rank, coords, part, lle_all in res:
    np.testing.assert_allclose(part, part_ref, rtol=1e-12)
    np.testing.assert_allclose(lle_all, lle_ref, rtol=1e-13)
