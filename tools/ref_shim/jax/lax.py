import numpy as _np

log = _np.log
exp = _np.exp
pow = _np.power
sqrt = _np.sqrt
abs = _np.abs
cos = _np.cos
sin = _np.sin


def cond(pred, t, f, *ops, operand=None):
    if not ops:
        ops = (operand,)
    return t(*ops) if pred else f(*ops)


def fori_loop(lo, hi, body, init):
    v = init
    for i in range(lo, hi):
        v = body(i, v)
    return v
