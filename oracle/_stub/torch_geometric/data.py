"""TEST INFRASTRUCTURE ONLY: attribute-bag stand-ins for torch_geometric.data.{Data,Batch}
(see oracle/_stub/torch_geometric/__init__.py)."""


class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


class Batch(Data):
    pass
