"""TEST INFRASTRUCTURE ONLY -- golden vectors of the reference's Stokes operators.

Runs the UNMODIFIED `swirl_fem/navier_stokes/navier_stokes.py` under the numpy
stubs of `oracle/ref_harness.py` on a tiny mesh (2 x 2 quads, periodic in y,
curved in x, velocity order 3 / pressure on 2 Gauss-Legendre points) and stores
every operator's action on seeded random fields in
`tests/golden/navier_stokes.npz`.  The harness transposes by unit-vector
probing inside python-loop vmaps, so this takes several minutes; it runs only
in the build container (`python -m oracle.make_golden_ns`).
"""

from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402


def main():
  ref = ref_harness.reference()
  from swirl_fem.core.premesh import Premesh  # pylint: disable=g-import-not-at-top
  from swirl_fem.navier_stokes import navier_stokes as ns  # pylint: disable=g-import-not-at-top
  out = {}
  for k in (1, 2, 3, 4):
    out[f'bdf{k}'] = np.asarray(ns.bdfk_coeffs(k))
  for k in (1, 2, 3):
    out[f'ext{k}'] = np.asarray(ns.extk_coeffs(k))

  ne, order = 2, 3
  pm = ref.premesh_commons.unit_cube_mesh(ne, ndim=2, periodic_dims=(1,))
  coords = np.asarray(pm.node_coords, dtype=np.float64)
  coords = np.stack([2 * coords[:, 0] - 1,
                     2 * np.pi * coords[:, 1] - np.pi], -1)
  # curved in x, still periodic in y
  coords[:, 0] += 0.1 * np.sin(coords[:, 1]) * (1 - coords[:, 0] ** 2)
  pm = Premesh.create(node_coords=coords, elements=np.asarray(pm.elements),
                      physical_groups=pm.physical_groups,
                      periodic_links=pm.periodic_links)
  sem = ns.StokesSEM.create(
      pm, boundary_conditions={'boundary': (ns.BCType.DIRICHLET, 0.0)},
      order=order)
  vmesh, pmesh = sem.velocity.vspace.mesh, sem.pressure.pspace.mesh
  out.update(
      ne=ne, order=order, premesh_coords=coords,
      v_coords=np.asarray(vmesh.node_coords),
      v_elements=np.asarray(vmesh.elements),
      p_coords=np.asarray(pmesh.node_coords),
      p_elements=np.asarray(pmesh.elements),
      interior_mask=np.asarray(sem.velocity.interior_mask),
      diag_qqt=np.asarray(sem.velocity.diag_qqt),
      velocity_mass_diag=np.asarray(sem.velocity_mass_diag))
  rng = np.random.default_rng(2024)
  u = rng.standard_normal((vmesh.num_nodes, 2))
  p = rng.standard_normal(pmesh.num_nodes)
  out.update(u=u, p=p)
  dt, k = 1e-3, 3
  ops = [
      ('A', lambda: sem.A(u)), ('B', lambda: sem.B(u)),
      ('Bi', lambda: sem.Bi(u)), ('C', lambda: sem.C(u)),
      ('D', lambda: sem.D(u)), ('Dt', lambda: sem.Dt(p)),
      ('Q', lambda: sem.Q(u, dt=dt, time_order=k)),
      ('E', lambda: sem.E(p, dt=dt, time_order=k)),
      ('filter', lambda: sem.filter(u, alpha=0.05)),
      ('vorticity', lambda: sem.vorticity(u)),
      ('pressure_B', lambda: sem.pressure.B(p)),
      ('project', lambda: ns._pressure_project_out_nullspace(sem, p)),  # pylint: disable=protected-access
      ('A_local', lambda: sem.velocity.A_local(sem.velocity.gather(u))),
      ('D_local', lambda: sem.D_local(sem.velocity.gather(u))),
      ('Dt_local', lambda: sem.Dt_local(sem.pressure.gather(p))),
  ]
  for name, fn in ops:
    t0 = time.time()
    out[name] = np.asarray(fn(), dtype=np.float64)
    print(f'{name}: {out[name].shape} in {time.time() - t0:.1f} s', flush=True)
  path = os.path.join(ROOT, 'tests', 'golden', 'navier_stokes.npz')
  np.savez_compressed(path, **out)
  print('wrote', path)


if __name__ == '__main__':
  main()
