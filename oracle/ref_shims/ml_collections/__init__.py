"""Stand-in for ml_collections.ConfigDict (attribute-access dict), used by configs/*.py."""


class ConfigDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v
