def make_axes_locatable(*a, **k):
    raise NotImplementedError
