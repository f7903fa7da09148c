"""Developer script: where does the lazy zero fill differ from the eager one?"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swirl_fem_b200 import _lib  # noqa: E402
from swirl_fem_b200.core.fespace import FiniteElementSpace  # noqa: E402
from swirl_fem_b200.core.interpolation import NodeType, Quadrature1D  # noqa: E402
from tests import helpers  # noqa: E402

GLL = NodeType.GAUSS_LOBATTO_LEGENDRE
piece = int(os.environ.get('PIECE', 1024))
with_mass = bool(int(os.environ.get('MASS', 1)))
ne = int(os.environ.get('NE', 16))
refined = helpers.deformed_premesh(3, ne, 8, seed=None, reorient=False)
mesh = refined.finalize(dtype=torch.float64)
space = FiniteElementSpace.create(mesh, Quadrature1D.create(8, GLL))
bmask = refined.finalize_host()['physical_masks']['boundary']
op = space.operator(dirichlet_mask=bmask, with_mass=with_mass)
dev = mesh.device
x = torch.randn(mesh.num_nodes, dtype=torch.float64, device=dev)
lam = 0.7 if with_mass else 0.0
y_e = op.apply(x, lam=lam, mu=1.3).clone()
ok = op.enable_lazy_zero(piece=piece)
print('enabled', ok, 'pieces', op._lazy[0].shape, 'chunks', op._lazy[1].numel() - 1)
pieces, cp = op._lazy[0].cpu().numpy(), op._lazy[1].cpu().numpy()
nz = int(_lib.lib().sfem_op_num_zero(op.handle))
for rep in range(3):
  out = torch.full_like(x, float('nan'))
  y_l = op.apply(x, lam=lam, mu=1.3, out=out)
  torch.cuda.synchronize()
  d = (y_l - y_e).abs()
  bad = torch.nonzero(~(d <= 1e-12 * float(y_e.abs().max()))).reshape(-1).cpu().numpy()
  print(f'rep {rep}: wrong dofs {bad.size} of {x.numel()} (prefix {nz}); timed out: {op.lazy_zero_timed_out()!r}')
  if bad.size:
    start, ln, ch = pieces[:, 0], pieces[:, 1] & 0xfff, pieces[:, 1] >> 12
    order = np.argsort(start)
    pos = np.searchsorted(start[order], bad, side='right') - 1
    pi = order[pos]
    off = bad - start[pi]
    yl, ye = y_l.cpu().numpy(), y_e.cpu().numpy()
    print('  in prefix:', int((bad < nz).sum()), ' nan:', int(np.isnan(yl[bad]).sum()),
          ' y_l == 0:', int((yl[bad] == 0).sum()))
    print('  chunks of the wrong dofs:', np.unique(ch[pi], return_counts=True))
    print('  offset within piece (min/median/max):', off.min(), np.median(off), off.max(),
          ' piece lens:', np.unique(ln[pi])[:10])
    print('  offset mod 32 histogram:', np.bincount(off % 32, minlength=32))
    for k in range(min(8, bad.size)):
      print(f'   dof {bad[k]} piece {pi[k]} start {start[pi[k]]} len {ln[pi[k]]} chunk {ch[pi[k]]} off {off[k]}  y_l {yl[bad[k]]:.6g} y_e {ye[bad[k]]:.6g}')
    # multiplicity of the wrong dofs
    el = mesh.elements.reshape(-1).cpu().numpy()
    cnt = np.bincount(el, minlength=x.numel())
    print('  multiplicity of wrong dofs:', np.unique(cnt[bad], return_counts=True))
