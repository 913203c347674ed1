class SamplePoints:
    def __init__(self, *a, **k):
        raise NotImplementedError("SamplePoints stub")
