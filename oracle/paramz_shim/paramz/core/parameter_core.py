"""TEST-ONLY paramz.core.parameter_core stand-in."""


class Parameterizable(object):
    def __init__(self, name=None, *a, **kw):
        self.name = name
        self._index_operations = {}
        self._parent_ = None

    def add_index_operation(self, name, operations):
        self._index_operations[name] = operations
        setattr(self, name, operations)

    def constrain_positive(self, warning=True):
        pass

    def constrain_negative(self, warning=True):
        pass

    def _add_to_index_operations(self, *a, **kw):
        pass

    def _remove_from_index_operations(self, *a, **kw):
        return []
