#!/usr/bin/env python
"""Poor man's pyflakes (no linter in the image): reports names that are read
but never bound in any enclosing scope of a Python source file.

  python tools/check_names.py swirl_fem_b200/navier_stokes/navier_stokes.py ...
"""
import ast
import builtins
import sys


class Scope:

  def __init__(self, parent=None):
    self.parent, self.bound = parent, set()

  def has(self, name):
    s = self
    while s is not None:
      if name in s.bound:
        return True
      s = s.parent
    return False


def bind_targets(node, scope):
  for n in ast.walk(node):
    if isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
      scope.bound.add(n.id)


def collect(node, scope):
  """Binds every name assigned anywhere in this scope's body (not nested)."""
  for child in ast.iter_child_nodes(node):
    if isinstance(child, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
      scope.bound.add(child.name)
      continue
    if isinstance(child, ast.Lambda):
      continue
    if isinstance(child, (ast.Import, ast.ImportFrom)):
      for a in child.names:
        scope.bound.add((a.asname or a.name).split('.')[0])
    if isinstance(child, ast.Name) and isinstance(child.ctx, ast.Store):
      scope.bound.add(child.id)
    if isinstance(child, ast.ExceptHandler) and child.name:
      scope.bound.add(child.name)
    if isinstance(child, (ast.ListComp, ast.SetComp, ast.DictComp,
                          ast.GeneratorExp)):
      continue
    collect(child, scope)


def check(node, scope, errors):
  if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
    inner = Scope(scope)
    a = node.args
    for arg in a.posonlyargs + a.args + a.kwonlyargs + [a.vararg, a.kwarg]:
      if arg is not None:
        inner.bound.add(arg.arg)
    for d in a.defaults + [d for d in a.kw_defaults if d is not None]:
      check(d, scope, errors)
    body = node.body if isinstance(node.body, list) else [node.body]
    for b in body:
      collect(ast.Module(body=[b], type_ignores=[]), inner)
    for b in body:
      check(b, inner, errors)
    return
  if isinstance(node, ast.ClassDef):
    inner = Scope(scope)
    collect(node, inner)
    for b in node.body:
      check(b, inner, errors)
    return
  if isinstance(node, (ast.ListComp, ast.SetComp, ast.DictComp,
                       ast.GeneratorExp)):
    inner = Scope(scope)
    for g in node.generators:
      bind_targets(g.target, inner)
    for child in ast.iter_child_nodes(node):
      check(child, inner, errors)
    return
  if isinstance(node, ast.Name) and isinstance(node.ctx, ast.Load):
    if (not scope.has(node.id) and not hasattr(builtins, node.id)
        and node.id not in ('__file__', '__name__', '__doc__')):
      errors.append((node.lineno, node.id))
  for child in ast.iter_child_nodes(node):
    check(child, scope, errors)


def main():
  bad = 0
  for path in sys.argv[1:]:
    tree = ast.parse(open(path).read(), path)
    top = Scope()
    collect(tree, top)
    errors = []
    for b in tree.body:
      check(b, top, errors)
    for line, name in sorted(set(errors)):
      print(f'{path}:{line}: undefined name {name!r}')
      bad += 1
  sys.exit(1 if bad else 0)


if __name__ == '__main__':
  main()
