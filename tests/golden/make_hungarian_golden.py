"""Extract the reference's own Hungarian test vectors into a committed fixture.

Reads /root/reference/hungarian_tf_tests.py (read-only, never imported: it needs
TensorFlow 0.12) with ``ast`` and evaluates only the ``np.array([...])`` literals, the
``p = 1e6`` / ``np.round(W * p) / p`` preprocessing and the expected ``*_t`` arrays.
Run in the build container only:  python tests/golden/make_hungarian_golden.py
Output: tests/golden/hungarian_kat.json
"""
import ast
import json
import os

import numpy as np

SRC = '/root/reference/hungarian_tf_tests.py'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hungarian_kat.json')


def main():
  tree = ast.parse(open(SRC).read())
  cases = []
  for node in ast.walk(tree):
    if not isinstance(node, ast.FunctionDef) or not node.name.startswith('test_'):
      continue
    env = {'np': np}
    for stmt in node.body:
      if isinstance(stmt, ast.Assign) and len(stmt.targets) == 1 and isinstance(stmt.targets[0], ast.Name):
        name = stmt.targets[0].id
        if name in ('W', 'p', 'c_0_t', 'c_1_t', 'M_t'):
          env[name] = eval(compile(ast.Expression(stmt.value), SRC, 'eval'), env)
    W = np.asarray(env['W'], dtype=np.float64)
    case = {
        'name': node.name,
        'line': node.lineno,
        'W': W.tolist(),
        # float32 is what the op sees (REGISTER_OP input "weights: float", hungarian.cc:27)
        'W_f32_hex': np.asarray(W, np.float32).tobytes().hex(),
        'shape': list(W.shape),
    }
    for k_src, k_dst in (('M_t', 'M'), ('c_0_t', 'cover_x'), ('c_1_t', 'cover_y')):
      if k_src in env:
        case[k_dst] = np.asarray(env[k_src], np.float64).tolist()
    case['kind'] = 'known_answer' if 'M' in case else 'termination_only'
    cases.append(case)
  cases.sort(key=lambda c: c['line'])
  with open(OUT, 'w') as f:
    json.dump({'source': 'renmengye/rec-attend-public hungarian_tf_tests.py', 'cases': cases}, f, indent=1)
  print('wrote', OUT, [(c['name'], c['kind'], c['shape']) for c in cases])


if __name__ == '__main__':
  main()
