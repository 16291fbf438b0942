class _Registry:
    def register_model(self, name):
        def deco(cls):
            return cls
        return deco


registry = _Registry()
