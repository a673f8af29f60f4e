def broadcast_shape(a, b):
    import torch
    return tuple(torch.broadcast_shapes(tuple(a), tuple(b)))
