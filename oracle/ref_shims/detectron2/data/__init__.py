class _Meta(dict):
    def __getattr__(self, k):
        return self.get(k)


class _Catalog:
    def get(self, name):
        return _Meta(name=name)


MetadataCatalog = _Catalog()
