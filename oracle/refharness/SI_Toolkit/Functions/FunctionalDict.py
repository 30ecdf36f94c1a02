"""TEST-ONLY shim of SI_Toolkit.Functions.FunctionalDict (reference Cost_Functions/__init__.py:103-104)."""


class FunctionalDict(dict):
    pass
