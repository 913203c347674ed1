def read_point_cloud(*a, **k):
    raise NotImplementedError("open3d stub")
