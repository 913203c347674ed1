class Data:
    """Attribute bag with the two derived counts the reference reads."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.pos.size(0)

    @property
    def num_edges(self):
        return self.edge_index.size(1)


class Batch(Data):
    pass


class InMemoryDataset:
    pass
