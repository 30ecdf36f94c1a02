"""TEST-ONLY shim of SI_Toolkit.General.variable_parameters.VariableParameters
(used at reference Controllers/__init__.py:85-86,106-107)."""


class VariableParameters:
    def __init__(self, lib=None):
        self.lib = lib

    def set_attributes(self, attributes, device=None):
        for k, v in attributes.items():
            setattr(self, k, v)

    def update_attributes(self, attributes):
        for k, v in attributes.items():
            if hasattr(self, k) and self.lib is not None and hasattr(getattr(self, k), "copy_"):
                self.lib.assign(getattr(self, k), self.lib.to_tensor(v, self.lib.float32))
            else:
                setattr(self, k, v)
