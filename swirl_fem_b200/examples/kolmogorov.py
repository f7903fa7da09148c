"""Kolmogorov-flow Navier-Stokes time stepping on the B200 path.

The reference's data generator (`swirl_fem/niles/datagen/datagen.py:54-197`)
advances the incompressible Navier-Stokes equations on the doubly periodic unit
square with a sinusoidal body force and linear drag: the convection term is
extrapolated (EXT-(k-1) of the stored `C(u)` history, :90-94), the Stokes part
is `StokesSEM.stokes_one_step` (BDF-k, :96-101).  This module is that loop on
`swirl_fem_b200.navier_stokes.StokesSEM` (every operator a CUDA kernel); the
HDF5 writing, flags and logging of the reference are not part of the path --
`one_cycle` returns the sampled states instead.
"""

from __future__ import annotations

import numpy as np
import torch

from swirl_fem_b200.common import premesh_commons
from swirl_fem_b200.navier_stokes import navier_stokes

# pylint: disable=invalid-name

# datagen.py:46-53
RESOLUTION = 64
ORDER = 8
TIME_ORDER = 3
REYNOLDS_NUMBER = 20000
DT = 1e-4
DRAG_COEFF = 0.1


def u_init_fn(x: torch.Tensor) -> torch.Tensor:
  """Initial velocity of the Kolmogorov flow at points `(G, 2)` (datagen.py:56-62)."""
  l = 2.
  u0 = torch.cos(2 * l * np.pi * x[:, 0]) * torch.sin(2 * l * np.pi * x[:, 1])
  u1 = -torch.sin(2 * l * np.pi * x[:, 0]) * torch.cos(2 * l * np.pi * x[:, 1])
  return torch.stack([u0, u1], dim=-1)


def forcing(x: torch.Tensor, u: torch.Tensor,
            drag_coeff: float = DRAG_COEFF) -> torch.Tensor:
  """Kolmogorov forcing with linear drag (datagen.py:65-72)."""
  k = 4.
  f0 = torch.sin(2 * np.pi * k * x[:, 1])
  return torch.stack([f0, torch.zeros_like(f0)], dim=-1) - drag_coeff * u


def compute_dx(mesh) -> float:
  """Minimum distance between the nodes of an element (datagen.py:75-85)."""
  x = mesh.element_coords().to(torch.float32)          # (E, n, d)
  n = x.shape[1]
  dx = np.inf
  eye = torch.diag(torch.full((n,), float('inf'), device=x.device))
  for start in range(0, x.shape[0], 4096):
    xe = x[start:start + 4096]
    pairwise = torch.linalg.norm(xe[:, :, None, :] - xe[:, None, :, :], dim=-1)
    dx = min(dx, float((pairwise + eye).min()))
  return dx


def solve_one_step(sem: navier_stokes.StokesSEM, us, ps, Cus, *,
                   reynolds_number: float = REYNOLDS_NUMBER, dt: float = DT,
                   time_order: int = TIME_ORDER, drag_coeff: float = DRAG_COEFF,
                   tol: float = 1e-5, atol: float = 1e-4):
  """One Navier-Stokes step (`_solve_one_step`, datagen.py:88-102).

  Returns `(u, p, C(u), aux)`; `aux` holds the two CG infos of the Stokes step.
  """
  # Extrapolate advection term
  ext_coeffs = navier_stokes.extk_coeffs(k=time_order - 1)
  Cu = sum(float(ext_coeffs[-i]) * Cus[-i]
           for i in range(1, len(ext_coeffs) + 1))
  # Solve the stokes system with extrapolated advection term
  f = forcing(sem.velocity.mesh.node_coords, us[-1], drag_coeff)
  f = -Cu + sem.B(f)
  u, p, aux = sem.stokes_one_step(us, ps, f, mu=1 / reynolds_number, dt=dt,
                                  time_order=time_order, tol=tol, atol=atol)
  return u, p, sem.C(u), aux


def one_cycle(sem: navier_stokes.StokesSEM, start_step: int, num_steps: int,
              us, ps, sample_every: int = 10, **step_kwargs):
  """A simulation cycle (datagen.py:105-172) without the file output.

  Returns `(us, ps, dataset)` with `dataset = {'t', 'u', 'p'}` sampled every
  `sample_every` steps (the reference samples every 10).
  """
  dt = step_kwargs.get('dt', DT)
  t = start_step * dt
  dataset = {'t': [t], 'u': [us[-1]], 'p': [ps[-1]]}
  us, ps = tuple(us), tuple(ps)
  Cus = tuple(map(sem.C, us))
  for step_idx in range(1, num_steps + 1):
    t += dt
    u, p, Cu, _ = solve_one_step(sem, us, ps, Cus, **step_kwargs)
    us = us[1:] + (u,)
    ps = ps[1:] + (p,)
    Cus = Cus[1:] + (Cu,)
    if step_idx % sample_every == 0:
      dataset['t'].append(t)
      dataset['u'].append(u)
      dataset['p'].append(p)
  dataset = {'t': np.asarray(dataset['t']),
             'u': torch.stack(dataset['u']), 'p': torch.stack(dataset['p'])}
  return us, ps, dataset


def create_sem(resolution: int = RESOLUTION, order: int = ORDER, device=None,
               dtype=None) -> navier_stokes.StokesSEM:
  """Doubly periodic unit square, no boundary conditions (datagen.py:177-181)."""
  premesh = premesh_commons.unit_cube_mesh(resolution, ndim=2,
                                           periodic_dims=(0, 1))
  return navier_stokes.StokesSEM.create(premesh, boundary_conditions={},
                                        order=order, device=device, dtype=dtype)


def initial_state(sem: navier_stokes.StokesSEM, time_order: int = TIME_ORDER):
  """`TIME_ORDER` copies of the initial velocity and a zero pressure
  (datagen.py:187-191)."""
  u_init = u_init_fn(sem.velocity.mesh.node_coords)
  p_init = torch.zeros(sem.pressure.pspace.mesh.num_nodes,
                       dtype=u_init.dtype, device=u_init.device)
  return (u_init,) * time_order, (p_init,) * time_order


def run_simulation(resolution: int = RESOLUTION, order: int = ORDER,
                   num_cycles: int = 1, num_steps_per_cycle: int = 10,
                   **step_kwargs):
  """`run_simulation` (datagen.py:175-201); yields `(cycle, dataset, cfl)`."""
  sem = create_sem(resolution, order)
  mesh_dx = compute_dx(sem.velocity.mesh)
  us, ps = initial_state(sem, step_kwargs.get('time_order', TIME_ORDER))
  dt = step_kwargs.get('dt', DT)
  for cycle_idx in range(num_cycles):
    us, ps, dataset = one_cycle(
        sem, cycle_idx * num_steps_per_cycle, num_steps_per_cycle, us, ps,
        **step_kwargs)
    cfl = float(us[-1].max()) * dt / mesh_dx
    yield cycle_idx, dataset, cfl
