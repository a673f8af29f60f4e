def pfor(loop_fn, iters):
    raise NotImplementedError("pfor is outside the shim: golden vectors only cover single evaluations")
