def point_sample(*a, **k):
    raise NotImplementedError("training only")


def get_uncertain_point_coords_with_randomness(*a, **k):
    raise NotImplementedError("training only")
