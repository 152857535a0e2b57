"""Namespace registration of the polars stand-in (TEST INFRASTRUCTURE): `pl.api.register_lazyframe_namespace("pb")(cls)`."""
_NAMESPACES = {}


def register_lazyframe_namespace(name):
    def deco(cls):
        _NAMESPACES[("LazyFrame", name)] = cls
        return cls

    return deco


def register_dataframe_namespace(name):
    def deco(cls):
        _NAMESPACES[("DataFrame", name)] = cls
        return cls

    return deco
