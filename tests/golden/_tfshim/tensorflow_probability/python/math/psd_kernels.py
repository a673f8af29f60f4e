class _K(object): pass
ExponentiatedQuadratic = MaternOneHalf = _K
