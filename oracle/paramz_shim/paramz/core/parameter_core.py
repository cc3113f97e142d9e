"""TEST-ONLY paramz.core.parameter_core stand-in."""


class Parameterizable(object):
    def __init__(self, name=None, *a, **kw):
        self.name = name
        self._index_operations = {}
        self._parent_ = None

    def copy(self, memo=None):
        """paramz Parameterizable.copy: a deep copy that is not linked to the original's parent"""
        import copy as _copy
        parent, self._parent_ = self._parent_, None
        try:
            c = _copy.deepcopy(self)
        finally:
            self._parent_ = parent
        return c

    def __getstate__(self):
        return dict(self.__dict__)

    def __setstate__(self, state):
        self.__dict__.update(state)

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
