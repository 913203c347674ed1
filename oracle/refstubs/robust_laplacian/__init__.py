def point_cloud_laplacian(*a, **k):
    raise NotImplementedError("robust_laplacian stub")
