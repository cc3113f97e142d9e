"""TEST-ONLY paramz.core.index_operations stand-in."""


class ParameterIndexOperations(object):
    size = 0

    def items(self):
        return []

    def properties_for(self, idx):
        return []
