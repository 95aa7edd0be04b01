def NoGradient(op_type):  # modellib.py:11 registers "Hungarian" as non-differentiable
  return None
